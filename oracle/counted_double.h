/*
 * counted_double.h -- a drop-in scalar that behaves like `double` and counts every floating-point operation performed on it.
 * TEST / MEASUREMENT INFRASTRUCTURE (oracle/): hmp_oracle_count.cpp compiles the CPU oracle a second time with `double`
 * replaced by this type, which yields the INSTRUMENTED operation count per candidate-step that SURVEY.md 8d / BASELINE.md 3
 * ask for (the survey's W is an estimate read off the reference's statements). Results are bit-identical to the plain
 * build: every operator forwards to the same IEEE double operation.
 *
 * Categories (per thread): add (+, -), mul, div, sqrt, exp (exp, pow), trig (sin, cos, acos), atan2, cmp (<, <=, >, >=, ==,
 * !=, min/max through them, isnan), rnd (floor, ceil, casts to int / float). Negation, fabs and copies are free. hypot(a, b)
 * is booked as 2 mul + 1 add + 1 sqrt (what it computes); pow(a, b) as one exp-class operation.
 */
#ifndef HMP_COUNTED_DOUBLE_H_
#define HMP_COUNTED_DOUBLE_H_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>

namespace hmp_count {

enum Op { ADD = 0, MUL, DIV, SQRT, EXP, TRIG, ATAN2, CMP, RND, N_OPS };

struct Counters {
	uint64_t n[N_OPS];
};
inline Counters& counters() {
	static thread_local Counters c{};
	return c;
}
inline void tick(Op o, uint64_t k = 1) { counters().n[o] += k; }

struct cd {
	double v;
	constexpr cd() : v(0.0) {}
	constexpr cd(double x) : v(x) {}
	explicit operator double() const { return v; }
	explicit operator float() const { tick(RND); return (float)v; }
	explicit operator int() const { tick(RND); return (int)v; }
	explicit operator unsigned() const { tick(RND); return (unsigned)v; }
	explicit operator long() const { tick(RND); return (long)v; }
	explicit operator long long() const { tick(RND); return (long long)v; }
	explicit operator unsigned long() const { tick(RND); return (unsigned long)v; }
	explicit operator bool() const { return v != 0.0; }
	cd& operator+=(cd o) { tick(ADD); v += o.v; return *this; }
	cd& operator-=(cd o) { tick(ADD); v -= o.v; return *this; }
	cd& operator*=(cd o) { tick(MUL); v *= o.v; return *this; }
	cd& operator/=(cd o) { tick(DIV); v /= o.v; return *this; }
	cd operator-() const { return cd(-v); }
	cd operator+() const { return *this; }
};

#define HMP_CD_BINOP(op, cat)                                                            \
	inline cd operator op(cd a, cd b) { tick(cat); return cd(a.v op b.v); }              \
	inline cd operator op(cd a, double b) { tick(cat); return cd(a.v op b); }            \
	inline cd operator op(double a, cd b) { tick(cat); return cd(a op b.v); }
HMP_CD_BINOP(+, ADD)
HMP_CD_BINOP(-, ADD)
HMP_CD_BINOP(*, MUL)
HMP_CD_BINOP(/, DIV)
#undef HMP_CD_BINOP
#define HMP_CD_CMP(op)                                                                   \
	inline bool operator op(cd a, cd b) { tick(CMP); return a.v op b.v; }                \
	inline bool operator op(cd a, double b) { tick(CMP); return a.v op b; }              \
	inline bool operator op(double a, cd b) { tick(CMP); return a op b.v; }
HMP_CD_CMP(<)
HMP_CD_CMP(<=)
HMP_CD_CMP(>)
HMP_CD_CMP(>=)
HMP_CD_CMP(==)
HMP_CD_CMP(!=)
#undef HMP_CD_CMP

}  // namespace hmp_count

// <cmath> look-alikes; the oracle calls them qualified (std::sqrt ...), so they have to live in namespace std
namespace std {
using hmp_count::cd;
inline cd sqrt(cd a) { hmp_count::tick(hmp_count::SQRT); return cd(std::sqrt(a.v)); }
inline cd exp(cd a) { hmp_count::tick(hmp_count::EXP); return cd(std::exp(a.v)); }
inline cd pow(cd a, cd b) { hmp_count::tick(hmp_count::EXP); return cd(std::pow(a.v, b.v)); }
inline cd pow(cd a, int b) { hmp_count::tick(hmp_count::EXP); return cd(std::pow(a.v, b)); }
inline cd pow(cd a, double b) { hmp_count::tick(hmp_count::EXP); return cd(std::pow(a.v, b)); }
inline cd sin(cd a) { hmp_count::tick(hmp_count::TRIG); return cd(std::sin(a.v)); }
inline cd cos(cd a) { hmp_count::tick(hmp_count::TRIG); return cd(std::cos(a.v)); }
inline cd acos(cd a) { hmp_count::tick(hmp_count::TRIG); return cd(std::acos(a.v)); }
inline cd atan2(cd a, cd b) { hmp_count::tick(hmp_count::ATAN2); return cd(std::atan2(a.v, b.v)); }
inline cd atan2(cd a, double b) { hmp_count::tick(hmp_count::ATAN2); return cd(std::atan2(a.v, b)); }
inline cd atan2(double a, cd b) { hmp_count::tick(hmp_count::ATAN2); return cd(std::atan2(a, b.v)); }
inline cd hypot(cd a, cd b) {
	hmp_count::tick(hmp_count::MUL, 2);
	hmp_count::tick(hmp_count::ADD);
	hmp_count::tick(hmp_count::SQRT);
	return cd(std::hypot(a.v, b.v));
}
inline cd hypot(cd a, double b) { return hypot(a, cd(b)); }
inline cd hypot(double a, cd b) { return hypot(cd(a), b); }
inline cd min(cd a, double b) { return std::min(a, cd(b)); }
inline cd min(double a, cd b) { return std::min(cd(a), b); }
inline cd max(cd a, double b) { return std::max(a, cd(b)); }
inline cd max(double a, cd b) { return std::max(cd(a), b); }
inline cd abs(cd a) { return cd(std::fabs(a.v)); }
inline cd fabs(cd a) { return cd(std::fabs(a.v)); }
inline cd floor(cd a) { hmp_count::tick(hmp_count::RND); return cd(std::floor(a.v)); }
inline cd ceil(cd a) { hmp_count::tick(hmp_count::RND); return cd(std::ceil(a.v)); }
inline cd round(cd a) { hmp_count::tick(hmp_count::RND); return cd(std::round(a.v)); }
inline bool isnan(cd a) { hmp_count::tick(hmp_count::CMP); return std::isnan(a.v); }
inline bool isinf(cd a) { hmp_count::tick(hmp_count::CMP); return std::isinf(a.v); }
template <>
struct numeric_limits<hmp_count::cd> {
	static constexpr bool is_specialized = true;
	static constexpr cd max() { return cd(numeric_limits<double>::max()); }
	static constexpr cd min() { return cd(numeric_limits<double>::min()); }
	static constexpr cd lowest() { return cd(numeric_limits<double>::lowest()); }
	static constexpr cd infinity() { return cd(numeric_limits<double>::infinity()); }
	static constexpr cd quiet_NaN() { return cd(numeric_limits<double>::quiet_NaN()); }
	static constexpr cd epsilon() { return cd(numeric_limits<double>::epsilon()); }
};
}  // namespace std

#endif
