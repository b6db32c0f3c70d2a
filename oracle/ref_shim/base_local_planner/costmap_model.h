// Stand-in for base_local_planner/costmap_model.h. footprintCost forwards to the oracle's statement of
// CostmapModel::footprintCost (Bresenham edges, -1 lethal / -2 unknown / -3 off-map): third-party, not validated here.
#pragma once
#include <base_local_planner/world_model.h>
#include <costmap_2d/costmap_2d.h>
#include <shim_hooks.h>
namespace base_local_planner {
class CostmapModel : public WorldModel {
public:
	CostmapModel(const costmap_2d::Costmap2D& costmap) : costmap_(costmap) {}
	double footprintCost(double x, double y, double theta, const std::vector<geometry_msgs::Point>& footprint_spec,
	                     double /*inscribed_radius*/ = 0.0, double /*circumscribed_radius*/ = 0.0) override {
		std::vector<double> xy;
		for (const auto& p : footprint_spec) {
			xy.push_back(p.x);
			xy.push_back(p.y);
		}
		return orc_tp_footprint_cost(costmap_.getCharMap(), (int)costmap_.getSizeInCellsX(), (int)costmap_.getSizeInCellsY(),
		                             costmap_.getOriginX(), costmap_.getOriginY(), costmap_.getResolution(), x, y, theta, xy.data(),
		                             (int)footprint_spec.size());
	}
private:
	const costmap_2d::Costmap2D& costmap_;
};
}  // namespace base_local_planner
