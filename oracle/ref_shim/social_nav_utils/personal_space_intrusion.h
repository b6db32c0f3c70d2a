// Stand-in for social_nav_utils/personal_space_intrusion.h -> oracle formulation (third-party, not validated here).
#pragma once
#include <shim_hooks.h>
namespace social_nav_utils {
class PersonalSpaceIntrusion {
public:
	PersonalSpaceIntrusion(double xp, double yp, double yawp, double cxx, double cxy, double cyx, double cyy, double var_front,
	                       double var_rear, double var_side, double xr, double yr, bool /*unify_asymmetry*/ = false)
	    : scale_(orc_tp_personal_space(xp, yp, yawp, cxx, cxy, cyx, cyy, var_front, var_rear, var_side, xr, yr)) {}
	void normalize() {}
	double getScale() const { return scale_; }
private:
	double scale_;
};
}  // namespace social_nav_utils
