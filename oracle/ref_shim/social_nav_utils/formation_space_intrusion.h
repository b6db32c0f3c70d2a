// Stand-in for social_nav_utils/formation_space_intrusion.h -> oracle formulation (third-party, not validated here).
#pragma once
#include <shim_hooks.h>
namespace social_nav_utils {
class FormationSpaceIntrusion {
public:
	FormationSpaceIntrusion(double xg, double yg, double yawg, double var_x, double var_y, double cxx, double cxy, double cyy,
	                        double xr, double yr)
	    : scale_(orc_tp_formation_space(xg, yg, yawg, var_x, var_y, cxx, cxy, cyy, xr, yr)) {}
	void normalize() {}
	double getScale() const { return scale_; }
private:
	double scale_;
};
}  // namespace social_nav_utils
