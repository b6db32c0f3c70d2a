// Stand-in for social_nav_utils/passing_speed_comfort.h -> oracle formulation (third-party, not validated here).
#pragma once
#include <shim_hooks.h>
namespace social_nav_utils {
class PassingSpeedComfort {
public:
	PassingSpeedComfort(double distance, double speed, double min_dist, double max_speed)
	    : v_(orc_tp_passing_speed(distance, speed, min_dist, max_speed)) {}
	double getDiscomfortNormalized() const { return v_; }
private:
	double v_;
};
}  // namespace social_nav_utils
