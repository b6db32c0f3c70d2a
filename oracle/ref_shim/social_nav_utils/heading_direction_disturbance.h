// Stand-in for social_nav_utils/heading_direction_disturbance.h -> oracle formulation (third-party, not validated here).
#pragma once
#include <shim_hooks.h>
namespace social_nav_utils {
class HeadingDirectionDisturbance {
public:
	HeadingDirectionDisturbance(double xp, double yp, double yawp, double cxx, double cxy, double cyy, double xr, double yr,
	                            double yawr, double vxr, double vyr, double person_radius, double fov_person)
	    : a_{xp, yp, yawp, cxx, cxy, cyy, xr, yr, yawr, vxr, vyr, person_radius, fov_person}, scale_(0.0) {}
	void normalize(double robot_circumradius, double max_speed) {
		scale_ = orc_tp_heading_disturbance(a_[0], a_[1], a_[2], a_[3], a_[4], a_[5], a_[6], a_[7], a_[8], a_[9], a_[10], a_[11],
		                                    a_[12], robot_circumradius, max_speed);
	}
	double getScale() const { return scale_; }
private:
	double a_[13];
	double scale_;
};
}  // namespace social_nav_utils
