// Stand-in for ignition/math/Helpers.hh (ignition-math 4): constants and equal().
#pragma once
#include <cmath>
#include <limits>
#define IGN_PI 3.14159265358979323846
#define IGN_PI_2 1.57079632679489661923
#define IGN_PI_4 0.78539816339744830962
#define IGN_DTOR(d) ((d) * IGN_PI / 180)
#define IGN_RTOD(r) ((r) * 180 / IGN_PI)
namespace ignition {
namespace math {
template <typename T>
inline bool equal(const T& a, const T& b, const T& epsilon = T(1e-6)) {
	volatile T diff = std::abs(a - b);
	return diff <= epsilon;
}
}  // namespace math
}  // namespace ignition
