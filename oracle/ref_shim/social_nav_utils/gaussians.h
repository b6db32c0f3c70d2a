// Stand-in for social_nav_utils/gaussians.h -> the oracle's statement of the formulation (third-party, not validated here).
#pragma once
#include <shim_hooks.h>
namespace social_nav_utils {
inline double calculateGaussianAngle(double x, double mean, double variance, bool normalize = false) {
	return orc_tp_gaussian_angle(x, mean, variance, normalize ? 1 : 0);
}
}  // namespace social_nav_utils
