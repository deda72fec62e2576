// Exact time bookkeeping of the reference as value types for the C++ side of the
// drop-in boundary: Rational, Slab, Time, TimeDelta, TimeStepId
// with the reference's semantics (src/Time/Slab.hpp, Time.hpp, TimeStepId.hpp:22-62,
// src/Utilities/Rational.hpp) -- a Time is a slab plus an exact rational fraction of it,
// its double value is (1 - f) start + f end; a TimeStepId orders (slab number, step time,
// substep) and moves to the next slab when a step ends on the slab boundary.
//
// Own implementation (header only, no dependencies); the known answers of the reference's
// tests/Unit/Time/Test_{Slab,Time,TimeStepId}.cpp are re-checked in tests/helpers/
// time_types_test.cpp.  DgTimeLoop at the end drives libdgrhs.so with these types the way
// the reference's AdvanceTime action does (Time/Actions/AdvanceTime.hpp): next id from the
// time stepper's next_time_id, the substep time handed to the library is the id's; the
// library forms the same times itself after dgrhs_set_slab (checked bit for bit).
#pragma once

#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <numeric>
#include <ostream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/dgrhs.h"

namespace spectre_b200 {

// ---- Rational (Utilities/Rational.hpp): 32-bit numerator / denominator in lowest terms,
// denominator > 0; intermediate products in 64 bits, overflow is an error ---------------
class Rational {
 public:
  Rational() = default;
  Rational(std::int32_t numerator, std::int32_t denominator) { assign(numerator, denominator); }
  Rational(std::int32_t integer) : num_(integer) {}  // NOLINT: implicit like the reference's
  std::int32_t numerator() const { return num_; }
  std::int32_t denominator() const { return den_; }
  double value() const { return static_cast<double>(num_) / static_cast<double>(den_); }
  Rational inverse() const {
    if (num_ == 0) throw std::domain_error("Rational: division by zero");
    return {den_, num_};
  }
  Rational operator-() const { return raw(-static_cast<std::int64_t>(num_), den_); }
  Rational& operator+=(const Rational& o) {
    return *this = raw(static_cast<std::int64_t>(num_) * o.den_ + static_cast<std::int64_t>(o.num_) * den_,
                       static_cast<std::int64_t>(den_) * o.den_);
  }
  Rational& operator-=(const Rational& o) { return *this += -o; }
  Rational& operator*=(const Rational& o) {
    return *this = raw(static_cast<std::int64_t>(num_) * o.num_, static_cast<std::int64_t>(den_) * o.den_);
  }
  Rational& operator/=(const Rational& o) { return *this *= o.inverse(); }
  friend Rational operator+(Rational a, const Rational& b) { return a += b; }
  friend Rational operator-(Rational a, const Rational& b) { return a -= b; }
  friend Rational operator*(Rational a, const Rational& b) { return a *= b; }
  friend Rational operator/(Rational a, const Rational& b) { return a /= b; }
  friend bool operator==(const Rational& a, const Rational& b) { return a.num_ == b.num_ && a.den_ == b.den_; }
  friend bool operator!=(const Rational& a, const Rational& b) { return !(a == b); }
  friend bool operator<(const Rational& a, const Rational& b) {
    return static_cast<std::int64_t>(a.num_) * b.den_ < static_cast<std::int64_t>(b.num_) * a.den_;
  }
  friend bool operator>(const Rational& a, const Rational& b) { return b < a; }
  friend bool operator<=(const Rational& a, const Rational& b) { return !(b < a); }
  friend bool operator>=(const Rational& a, const Rational& b) { return !(a < b); }
  friend std::ostream& operator<<(std::ostream& os, const Rational& r) {
    return os << r.num_ << '/' << r.den_;
  }

 private:
  static Rational raw(std::int64_t n, std::int64_t d) {
    if (d == 0) throw std::domain_error("Rational: zero denominator");
    if (d < 0) {
      n = -n;
      d = -d;
    }
    const std::int64_t g = std::gcd(n < 0 ? -n : n, d);
    if (g > 1) {
      n /= g;
      d /= g;
    }
    if (n > std::numeric_limits<std::int32_t>::max() || n < std::numeric_limits<std::int32_t>::min() ||
        d > std::numeric_limits<std::int32_t>::max())
      throw std::overflow_error("Rational: overflow");
    Rational r;
    r.num_ = static_cast<std::int32_t>(n);
    r.den_ = static_cast<std::int32_t>(d);
    return r;
  }
  void assign(std::int64_t n, std::int64_t d) { *this = raw(n, d); }
  std::int32_t num_ = 0, den_ = 1;
};

class Time;
class TimeDelta;

// ---- Slab (Time/Slab.hpp): a [start, end] interval of doubles ----------------------------
class Slab {
 public:
  Slab() = default;
  Slab(double start, double end) : start_(start), end_(end) {
    if (!(start_ < end_)) throw std::invalid_argument("Slab: backwards or empty");
  }
  static Slab with_duration_from_start(double start, double duration) { return {start, start + duration}; }
  static Slab with_duration_to_end(double end, double duration) { return {end - duration, end}; }
  inline Time start() const;
  inline Time end() const;
  inline TimeDelta duration() const;
  Slab advance() const { return {end_, end_ + (end_ - start_)}; }
  Slab retreat() const { return {start_ - (end_ - start_), start_}; }
  inline Slab advance_towards(const TimeDelta& dt) const;
  Slab with_duration_from_start(double duration) const { return {start_, start_ + duration}; }
  Slab with_duration_to_end(double duration) const { return {end_ - duration, end_}; }
  bool is_followed_by(const Slab& other) const { return end_ == other.start_; }
  bool is_preceeded_by(const Slab& other) const { return other.is_followed_by(*this); }
  bool overlaps(const Slab& other) const { return !(end_ <= other.start_ || start_ >= other.end_); }
  friend bool operator==(const Slab& a, const Slab& b) { return a.start_ == b.start_ && a.end_ == b.end_; }
  friend bool operator!=(const Slab& a, const Slab& b) { return !(a == b); }
  friend bool operator<(const Slab& a, const Slab& b) {
    if (!(a == b || a.end_ <= b.start_ || a.start_ >= b.end_))
      throw std::logic_error("cannot compare overlapping slabs");
    return a.end_ <= b.start_;
  }
  friend bool operator>(const Slab& a, const Slab& b) { return b < a; }
  friend bool operator<=(const Slab& a, const Slab& b) { return !(a > b); }
  friend bool operator>=(const Slab& a, const Slab& b) { return !(a < b); }
  double start_value() const { return start_; }
  double end_value() const { return end_; }

 private:
  double start_ = std::numeric_limits<double>::quiet_NaN();
  double end_ = std::numeric_limits<double>::quiet_NaN();
};

// ---- TimeDelta (Time/Time.hpp): a signed rational fraction of a slab ----------------------
class TimeDelta {
 public:
  using rational_t = Rational;
  TimeDelta() = default;
  TimeDelta(Slab slab, rational_t fraction) : slab_(slab), fraction_(fraction) {}
  TimeDelta with_slab(const Slab& new_slab) const { return {new_slab, fraction_}; }
  const Slab& slab() const { return slab_; }
  const rational_t& fraction() const { return fraction_; }
  double value() const { return (slab_.end_value() - slab_.start_value()) * fraction_.value(); }
  bool is_positive() const { return fraction_ > 0; }
  TimeDelta& operator+=(const TimeDelta& o) {
    same_slab(o);
    fraction_ += o.fraction_;
    return *this;
  }
  TimeDelta& operator-=(const TimeDelta& o) {
    same_slab(o);
    fraction_ -= o.fraction_;
    return *this;
  }
  TimeDelta operator+() const { return *this; }
  TimeDelta operator-() const { return {slab_, -fraction_}; }
  TimeDelta& operator*=(const rational_t& m) {
    fraction_ *= m;
    return *this;
  }
  TimeDelta& operator/=(const rational_t& d) {
    fraction_ /= d;
    return *this;
  }
  friend TimeDelta operator+(TimeDelta a, const TimeDelta& b) { return a += b; }
  friend TimeDelta operator-(TimeDelta a, const TimeDelta& b) { return a -= b; }
  friend TimeDelta operator*(TimeDelta a, const rational_t& b) { return a *= b; }
  friend TimeDelta operator*(const rational_t& a, TimeDelta b) { return b *= a; }
  friend TimeDelta operator/(TimeDelta a, const rational_t& b) { return a /= b; }
  friend double operator/(const TimeDelta& a, const TimeDelta& b) {
    return (a.fraction_ / b.fraction_).value() *
           ((a.slab_.end_value() - a.slab_.start_value()) / (b.slab_.end_value() - b.slab_.start_value()));
  }
  friend bool operator==(const TimeDelta& a, const TimeDelta& b) {
    return a.slab_ == b.slab_ && a.fraction_ == b.fraction_;
  }
  friend bool operator!=(const TimeDelta& a, const TimeDelta& b) { return !(a == b); }
  friend bool operator<(const TimeDelta& a, const TimeDelta& b) {
    a.same_slab(b);
    return a.fraction_ < b.fraction_;
  }
  friend bool operator>(const TimeDelta& a, const TimeDelta& b) { return b < a; }
  friend bool operator<=(const TimeDelta& a, const TimeDelta& b) { return !(b < a); }
  friend bool operator>=(const TimeDelta& a, const TimeDelta& b) { return !(a < b); }

 private:
  void same_slab(const TimeDelta& o) const {
    if (slab_ != o.slab_) throw std::logic_error("TimeDeltas of different slabs");
  }
  Slab slab_;
  rational_t fraction_;
};

inline TimeDelta abs(TimeDelta t) { return t.is_positive() ? t : -t; }

// ---- Time (Time/Time.hpp): a slab and an exact fraction in [0, 1] of it -------------------
class Time {
 public:
  using rational_t = Rational;
  Time() = default;
  Time(Slab slab, rational_t fraction) : slab_(slab), fraction_(fraction) {
    range_check();
    compute_value();
  }
  // the same instant expressed in an adjacent (or identical) slab; only slab boundaries move
  Time with_slab(const Slab& new_slab) const {
    if (new_slab == slab_) return *this;
    if (is_at_slab_start()) {
      if (slab_.start_value() == new_slab.start_value()) return new_slab.start();
      if (slab_.start_value() != new_slab.end_value()) throw std::logic_error("Time: cannot move to that slab");
      return new_slab.end();
    }
    if (!is_at_slab_end()) throw std::logic_error("Time: only slab boundaries can change slab");
    if (slab_.end_value() == new_slab.end_value()) return new_slab.end();
    if (slab_.end_value() != new_slab.start_value()) throw std::logic_error("Time: cannot move to that slab");
    return new_slab.start();
  }
  double value() const { return value_; }
  const Slab& slab() const { return slab_; }
  const rational_t& fraction() const { return fraction_; }
  Time& operator+=(const TimeDelta& d) {
    *this = with_slab(d.slab());
    fraction_ += d.fraction();
    range_check();
    compute_value();
    return *this;
  }
  Time& operator-=(const TimeDelta& d) { return *this += -d; }
  bool is_at_slab_start() const { return fraction_ == 0; }
  bool is_at_slab_end() const { return fraction_ == 1; }
  bool is_at_slab_boundary() const { return is_at_slab_start() || is_at_slab_end(); }
  friend Time operator+(Time a, const TimeDelta& b) { return a += b; }
  friend Time operator+(const TimeDelta& a, Time b) { return b += a; }
  friend Time operator-(Time a, const TimeDelta& b) { return a -= b; }
  friend TimeDelta operator-(const Time& a, const Time& b) {
    if (a.slab_ == b.slab_) return {a.slab_, a.fraction_ - b.fraction_};
    // adjacent slabs: one of the two times must sit on the common boundary
    if (a.slab_.is_followed_by(b.slab_)) {
      if (a.is_at_slab_end()) return {b.slab_, -b.fraction_};
      if (!b.is_at_slab_start()) throw std::logic_error("cannot subtract times of different slabs");
      return {a.slab_, a.fraction_ - 1};
    }
    if (!a.slab_.is_preceeded_by(b.slab_)) throw std::logic_error("cannot subtract times of different slabs");
    if (a.is_at_slab_start()) return {b.slab_, Rational(1) - b.fraction_};
    if (!b.is_at_slab_end()) throw std::logic_error("cannot subtract times of different slabs");
    return {a.slab_, a.fraction_};
  }
  // the same instant in adjacent slabs compares equal (end of one == start of the next)
  friend bool operator==(const Time& a, const Time& b) {
    if (a.slab_ == b.slab_) return a.fraction_ == b.fraction_;
    return (a.is_at_slab_end() && b.is_at_slab_start() && a.slab_.end_value() == b.slab_.start_value()) ||
           (a.is_at_slab_start() && b.is_at_slab_end() && a.slab_.start_value() == b.slab_.end_value()) ||
           (a.is_at_slab_start() && b.is_at_slab_start() && a.slab_.start_value() == b.slab_.start_value()) ||
           (a.is_at_slab_end() && b.is_at_slab_end() && a.slab_.end_value() == b.slab_.end_value());
  }
  friend bool operator!=(const Time& a, const Time& b) { return !(a == b); }
  friend bool operator<(const Time& a, const Time& b) {
    if (a == b) return false;
    if (a.slab_ == b.slab_) return a.fraction_ < b.fraction_;
    return a.value_ < b.value_;
  }
  friend bool operator>(const Time& a, const Time& b) { return b < a; }
  friend bool operator<=(const Time& a, const Time& b) { return !(b < a); }
  friend bool operator>=(const Time& a, const Time& b) { return !(a < b); }

 private:
  void compute_value() {
    value_ = (Rational(1) - fraction_).value() * slab_.start_value() + fraction_.value() * slab_.end_value();
  }
  void range_check() const {
    if (fraction_ < 0 || fraction_ > 1) throw std::out_of_range("Time: slab fraction out of [0, 1]");
  }
  Slab slab_;
  rational_t fraction_;
  double value_ = std::numeric_limits<double>::quiet_NaN();
};

inline Time Slab::start() const { return {*this, 0}; }
inline Time Slab::end() const { return {*this, 1}; }
inline TimeDelta Slab::duration() const { return {*this, 1}; }
inline Slab Slab::advance_towards(const TimeDelta& dt) const {
  if (!dt.is_positive() && !(-dt).is_positive()) throw std::logic_error("cannot advance along a zero time vector");
  return dt.is_positive() ? advance() : retreat();
}

inline std::ostream& operator<<(std::ostream& os, const Slab& s) {
  return os << "Slab[" << s.start_value() << "," << s.end_value() << "]";
}
inline std::ostream& operator<<(std::ostream& os, const Time& t) {
  return os << t.slab() << ":" << t.fraction();
}
inline std::ostream& operator<<(std::ostream& os, const TimeDelta& d) {
  return os << d.slab() << ":" << d.fraction();
}

// ---- TimeStepId (Time/TimeStepId.hpp:22-62) -----------------------------------------------
class TimeStepId {
 public:
  TimeStepId() = default;
  // at the start of a step; a step that starts on the (evolution-direction) end of its slab
  // is moved to the next slab
  TimeStepId(bool time_runs_forward, std::int64_t slab_number, const Time& time)
      : slab_number_(slab_number), step_time_(time), step_size_(time_runs_forward ? 1 : -1),
        substep_time_(time.value()) {
    canonicalize();
  }
  // at substep `substep` (time `substep_time`) of the step that starts at `step_time`
  TimeStepId(bool time_runs_forward, std::int64_t slab_number, const Time& step_time, std::uint64_t substep,
             const TimeDelta& step_size, double substep_time)
      : slab_number_(slab_number), step_time_(step_time), substep_(static_cast<std::uint8_t>(substep)),
        step_size_(substep == 0 ? Rational(time_runs_forward ? 1 : -1) : step_size.fraction()),
        substep_time_(substep_time) {
    if (substep > std::numeric_limits<std::uint8_t>::max()) throw std::overflow_error("substep");
    if (substep_ == 0 && step_time_.value() != substep_time_)
      throw std::logic_error("initial substep must align with the step");
    if (substep_ != 0 && !(step_time.slab() == step_size.slab()))
      throw std::logic_error("time and step have different slabs");
    if (substep_ != 0 && time_runs_forward != step_size.is_positive())
      throw std::logic_error("step size has the wrong sign");
    canonicalize();
  }
  bool time_runs_forward() const { return step_size_ > 0; }
  std::int64_t slab_number() const { return slab_number_; }
  const Time& step_time() const { return step_time_; }
  std::uint64_t substep() const { return substep_; }
  TimeDelta step_size() const {
    if (substep_ == 0) throw std::logic_error("step size not available at substep 0");
    return {step_time_.slab(), step_size_};
  }
  double substep_time() const { return substep_time_; }
  bool is_at_slab_boundary() const { return substep_ == 0 && step_time_.is_at_slab_boundary(); }
  TimeStepId next_step(const TimeDelta& step_size) const {
    return {time_runs_forward(), slab_number_, step_time_ + step_size};
  }
  TimeStepId next_substep(const TimeDelta& step_size, double step_fraction) const {
    if (step_fraction < 0.0 || step_fraction > 1.0) throw std::out_of_range("substep must be within the step");
    const double new_time =
        (1.0 - step_fraction) * step_time_.value() + step_fraction * (step_time_ + step_size).value();
    return {time_runs_forward(), slab_number_, step_time_, substep() + 1, step_size, new_time};
  }
  friend bool operator==(const TimeStepId& a, const TimeStepId& b) {
    bool equal = a.slab_number_ == b.slab_number_ && a.step_time_ == b.step_time_ && a.substep_ == b.substep_;
    if (equal && a.substep_ != 0) equal = a.step_size() == b.step_size();
    return equal;
  }
  friend bool operator!=(const TimeStepId& a, const TimeStepId& b) { return !(a == b); }
  friend bool operator<(const TimeStepId& a, const TimeStepId& b) {
    if (a.slab_number_ != b.slab_number_) return a.slab_number_ < b.slab_number_;
    if (a.step_time_ != b.step_time_)
      return a.time_runs_forward() ? a.step_time_ < b.step_time_ : b.step_time_ < a.step_time_;
    if (a.substep_ != b.substep_) return a.substep_ < b.substep_;
    if (a.substep_ == 0) return false;
    return a.step_size() < b.step_size();
  }
  friend bool operator>(const TimeStepId& a, const TimeStepId& b) { return b < a; }
  friend bool operator<=(const TimeStepId& a, const TimeStepId& b) { return !(b < a); }
  friend bool operator>=(const TimeStepId& a, const TimeStepId& b) { return !(a < b); }
  friend std::ostream& operator<<(std::ostream& s, const TimeStepId& id) {
    return s << id.slab_number_ << ':' << id.step_time_ << ':' << static_cast<int>(id.substep_) << ':'
             << id.substep_time_;
  }

 private:
  void canonicalize() {
    if (time_runs_forward() ? step_time_.is_at_slab_end() : step_time_.is_at_slab_start()) {
      if (substep_ != 0) throw std::logic_error("time needs to be advanced, but the step already started");
      const Slab new_slab = time_runs_forward() ? step_time_.slab().advance() : step_time_.slab().retreat();
      ++slab_number_;
      step_time_ = step_time_.with_slab(new_slab);
    }
  }
  std::int64_t slab_number_ = std::numeric_limits<std::int64_t>::lowest();
  Time step_time_{};
  std::uint8_t substep_ = 0;
  Rational step_size_{};
  double substep_time_ = 0.0;
};

// ---- TimeStepper::next_time_id for the steppers of the path ------------------------------
// AdamsBashforth.cpp:92-96 (next_step), RungeKutta.cpp:34-58 (substep_times of the Butcher
// tableau), Rk3HesthavenSsp.cpp:36-48.  `stepper` is a DGRHS_STEPPER_* id, `order` is read
// for Adams-Bashforth only.
inline TimeStepId next_time_id(int stepper, int order, const TimeStepId& current_id, const TimeDelta& time_step) {
  int substeps = 0;
  if (dgrhs_stepper_properties(stepper, order, nullptr, &substeps, nullptr, nullptr))
    throw std::runtime_error(dgrhs_last_error());
  if (current_id.substep() >= static_cast<std::uint64_t>(substeps))
    throw std::logic_error("substep should be less than the number of substeps");
  if (current_id.substep() + 1 == static_cast<std::uint64_t>(substeps)) return current_id.next_step(time_step);
  std::vector<double> fractions(static_cast<size_t>(substeps));
  if (dgrhs_stepper_substep_fractions(stepper, fractions.data())) throw std::runtime_error(dgrhs_last_error());
  return current_id.next_substep(time_step, fractions[current_id.substep()]);
}

// ---- choose_lts_step_size (src/Time/ChooseLtsStepSize.cpp:14-39) -------------------------
// the step slab / 2^n closest to (not above) the desired step that still hits the slab boundary
// from `time`, a binary-fraction time of its slab; negative desired steps run backwards
inline TimeDelta choose_lts_step_size(const Time& time, double desired_step) {
  const auto den = time.fraction().denominator();
  if ((den & (den - 1)) != 0) throw std::logic_error("Not at a binary-fraction time within slab");
  const TimeDelta full_slab = desired_step > 0.0 ? time.slab().duration() : -time.slab().duration();
  const double desired_step_count = full_slab.value() / desired_step;
  // log2(2^n + eps) may give n: the inner ceil avoids that
  const std::size_t power =
      desired_step_count == 0.0 ? 0 : static_cast<std::size_t>(std::ceil(std::log2(std::ceil(desired_step_count))));
  const auto step_count = std::max(static_cast<decltype(den)>(1) << power, den);
  return full_slab / step_count;
}

// ---- TimeSteppers::adams_lts (src/Time/TimeSteppers/AdamsLts.hpp:27-139) -----------------
// lts_coefficients with the reference's argument meaning (explicit = Adams-Bashforth and
// implicit = Adams-Moulton schemes; a TimeStepId with substep() == 1 is the predictor value of
// its step): the ids of the local and the remote side of a mortar in the order of their
// insertion into the BoundaryHistory, the step [start_time, end_time] of the local side, the
// orders of the three schemes.  All times must lie in one slab (or in slabs of equal length
// that follow each other: they are brought to integer ticks of a common denominator).
namespace TimeSteppers::adams_lts {
enum class SchemeType { Explicit, Implicit };
struct AdamsScheme {
  SchemeType type;
  size_t order;
};
inline bool operator==(const AdamsScheme& a, const AdamsScheme& b) { return a.type == b.type && a.order == b.order; }
using LtsCoefficients = std::vector<std::tuple<TimeStepId, TimeStepId, double>>;

inline LtsCoefficients lts_coefficients(const std::vector<TimeStepId>& local_times,
                                        const std::vector<TimeStepId>& remote_times, const Time& start_time,
                                        const Time& end_time, const AdamsScheme& local_scheme,
                                        const AdamsScheme& remote_scheme, const AdamsScheme& small_step_scheme) {
  if (local_times.empty() || remote_times.empty()) throw std::runtime_error("adams_lts::lts_coefficients: empty history");
  // common tick: slabs counted from the slab of start_time, fractions over one denominator
  const Slab& base = start_time.slab();
  const double length = base.end_value() - base.start_value();
  std::int64_t den = 1;
  const auto slab_offset = [&](const Time& t) -> std::int64_t {
    const double k = (t.slab().start_value() - base.start_value()) / length;
    const auto r = static_cast<std::int64_t>(std::llround(k));
    if (std::abs(k - static_cast<double>(r)) > 1e-9 ||
        std::abs((t.slab().end_value() - t.slab().start_value()) - length) > 1e-9 * std::abs(length))
      throw std::runtime_error("adams_lts::lts_coefficients: slabs of different lengths");
    return r;
  };
  std::vector<const Time*> all{&start_time, &end_time};
  for (const auto& id : local_times) all.push_back(&id.step_time());
  for (const auto& id : remote_times) all.push_back(&id.step_time());
  for (const Time* t : all) den = std::lcm(den, static_cast<std::int64_t>(t->fraction().denominator()));
  const auto tick = [&](const Time& t) -> long long {
    return slab_offset(t) * den + static_cast<std::int64_t>(t.fraction().numerator()) *
                                      (den / static_cast<std::int64_t>(t.fraction().denominator()));
  };
  for (const auto* ids : {&local_times, &remote_times})
    for (const auto& id : *ids)
      if (id.substep() != 0)
        den = std::lcm(den, static_cast<std::int64_t>(id.step_size().fraction().denominator()));
  std::vector<long long> lt, rt, ls, rs;
  const auto substep_size = [&](const TimeStepId& id) -> long long {
    if (id.substep() == 0) return 0;
    if (id.substep() != 1) throw std::runtime_error("adams_lts::lts_coefficients: substep > 1");
    const auto f = id.step_size().fraction();
    return static_cast<std::int64_t>(f.numerator()) * (den / static_cast<std::int64_t>(f.denominator()));
  };
  for (const auto& id : local_times) {
    lt.push_back(tick(id.step_time()));
    ls.push_back(substep_size(id));
  }
  for (const auto& id : remote_times) {
    rt.push_back(tick(id.step_time()));
    rs.push_back(substep_size(id));
  }
  const int cap = 8 * 16;
  std::vector<int> li(cap), ri(cap);
  std::vector<double> cf(cap);
  int n = 0;
  const auto imp = [](const AdamsScheme& sch) { return sch.type == SchemeType::Implicit ? 1 : 0; };
  if (dgrhs_adams_lts_coefficients_general(
          imp(local_scheme), static_cast<int>(local_scheme.order), imp(remote_scheme),
          static_cast<int>(remote_scheme.order), imp(small_step_scheme),
          static_cast<int>(small_step_scheme.order), static_cast<int>(lt.size()), lt.data(), ls.data(),
          static_cast<int>(rt.size()), rt.data(), rs.data(), tick(start_time), tick(end_time),
          base.start_value(), length / static_cast<double>(den), cap, &n, li.data(), ri.data(),
          cf.data()) != 0)
    throw std::runtime_error(dgrhs_last_error());
  LtsCoefficients out;
  for (int t = 0; t < n; ++t) out.emplace_back(local_times[li[t]], remote_times[ri[t]], cf[t]);
  return out;
}
}  // namespace TimeSteppers::adams_lts

// ---- the time loop of the evolution executables on a resident batch ------------------------
// Global time stepping with a constant slab size and `steps_per_slab` equal steps per slab:
// the ids follow Time/Actions/AdvanceTime.hpp (id <- next id of the time stepper, the step is
// re-expressed in the new slab at a slab boundary), the library is put in slab mode
// (dgrhs_set_slab) and every RHS is evaluated at the id's substep_time(); the library's own
// substep time is checked against it bit for bit.  The self-start RHS evaluations
// (SelfStartActions.hpp, slab number -1 ... in the reference) are run first at the library's
// times.
class DgTimeLoop {
 public:
  DgTimeLoop(dgrhs_ctx* ctx, int stepper, int order, const Slab& initial_slab, int steps_per_slab)
      : ctx_(ctx), stepper_(stepper), order_(order),
        step_(initial_slab.duration() / Rational(steps_per_slab)),
        id_(true, 0, initial_slab.start()) {
    check(dgrhs_set_stepper(ctx_, stepper_, order_, initial_slab.start_value(), step_.value()));
    check(dgrhs_set_slab(ctx_, initial_slab.start_value(), initial_slab.end_value(), steps_per_slab));
  }
  const TimeStepId& time_step_id() const { return id_; }
  const TimeDelta& time_step() const { return step_; }
  // one RHS evaluation + substep update; returns true when a full step is done
  bool advance_substep() {
    int self_start = 0;
    check(dgrhs_self_start_substeps_left(ctx_, &self_start));
    double library_time = 0.0;
    check(dgrhs_begin_substep(ctx_, &library_time));
    double time = library_time;
    if (self_start == 0) {
      time = id_.substep_time();
      if (time != library_time) throw std::logic_error("substep time of the library differs from the TimeStepId's");
    }
    check(dgrhs_compute_time_derivative(ctx_, time, 0));
    int done = 0;
    check(dgrhs_end_substep(ctx_, &done));
    if (self_start == 0) {
      id_ = next_time_id(stepper_, order_, id_, step_);
      step_ = step_.with_slab(id_.step_time().slab());
      if (done != (id_.substep() == 0 ? 1 : 0)) throw std::logic_error("step boundary of the library and of the id differ");
    }
    return done != 0;
  }
  void take_steps(int n) {
    for (int k = 0; k < n; ++k)
      while (!advance_substep()) {
      }
  }

 private:
  static void check(int rc) {
    if (rc) throw std::runtime_error(dgrhs_last_error());
  }
  dgrhs_ctx* ctx_;
  int stepper_, order_;
  TimeDelta step_;
  TimeStepId id_;
};

}  // namespace spectre_b200

namespace std {
template <>
struct hash<spectre_b200::TimeStepId> {
  size_t operator()(const spectre_b200::TimeStepId& id) const {
    auto mix = [](size_t h, size_t v) { return h ^ (v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2)); };
    size_t h = std::hash<std::int64_t>{}(id.slab_number());
    h = mix(h, std::hash<double>{}(id.step_time().value()));
    h = mix(h, std::hash<std::int32_t>{}(id.step_time().fraction().numerator()));
    h = mix(h, std::hash<std::int32_t>{}(id.step_time().fraction().denominator()));
    h = mix(h, std::hash<std::uint64_t>{}(id.substep()));
    if (id.substep() != 0) {
      h = mix(h, std::hash<std::int32_t>{}(id.step_size().fraction().numerator()));
      h = mix(h, std::hash<std::int32_t>{}(id.step_size().fraction().denominator()));
      h = mix(h, std::hash<double>{}(id.substep_time()));
    }
    return h;
  }
};
}  // namespace std
