// Stand-in for base_local_planner/world_model.h: the 4-argument footprintCost convenience overload.
#pragma once
#include <geometry_msgs/Point.h>
#include <vector>
namespace base_local_planner {
class WorldModel {
public:
	virtual ~WorldModel() {}
	virtual double footprintCost(double x, double y, double theta, const std::vector<geometry_msgs::Point>& footprint_spec,
	                             double inscribed_radius = 0.0, double circumscribed_radius = 0.0) = 0;
};
}  // namespace base_local_planner
