// Stand-in for the UPSTREAM base_local_planner::MapGridCostFunction (path_costs_, goal_costs_ of HumapPlanner,
// include/humap_local_planner/humap_planner.h:641-676): same loop as the humap class without the neighbour heuristic.
// Third-party, restated from the published source; not validated by the _ref build.
#pragma once
#include <base_local_planner/map_grid.h>
#include <base_local_planner/trajectory_cost_function.h>
#include <cmath>
namespace base_local_planner {
enum CostAggregationType { Last, Sum, Product };
class MapGridCostFunction : public TrajectoryCostFunction {
public:
	MapGridCostFunction(costmap_2d::Costmap2D* costmap, double xshift = 0.0, double yshift = 0.0,
	                    bool is_local_goal_function = false, CostAggregationType aggregationType = Last)
	    : costmap_(costmap), map_(costmap->getSizeInCellsX(), costmap->getSizeInCellsY()), aggregationType_(aggregationType),
	      xshift_(xshift), yshift_(yshift), is_local_goal_function_(is_local_goal_function), stop_on_failure_(true) {}
	void setTargetPoses(std::vector<geometry_msgs::PoseStamped> target_poses) { target_poses_ = target_poses; }
	void setXShift(double xshift) { xshift_ = xshift; }
	void setYShift(double yshift) { yshift_ = yshift; }
	void setStopOnFailure(bool stop_on_failure) { stop_on_failure_ = stop_on_failure; }
	bool prepare() override {
		map_.resetPathDist();
		if (is_local_goal_function_) {
			map_.setLocalGoal(*costmap_, target_poses_);
		} else {
			map_.setTargetCells(*costmap_, target_poses_);
		}
		return true;
	}
	double obstacleCosts() { return map_.obstacleCosts(); }
	double unreachableCellCosts() { return map_.unreachableCellCosts(); }
	double getCellCosts(unsigned int px, unsigned int py) { return map_(px, py).target_dist; }
	double scoreTrajectory(Trajectory& traj) override {
		double cost = 0.0;
		if (aggregationType_ == Product) cost = 1.0;
		double px, py, pth;
		unsigned int cell_x, cell_y;
		for (unsigned int i = 0; i < traj.getPointsSize(); ++i) {
			traj.getPoint(i, px, py, pth);
			if (xshift_ != 0.0) {
				px = px + xshift_ * std::cos(pth);
				py = py + xshift_ * std::sin(pth);
			}
			if (yshift_ != 0.0) {
				px = px + yshift_ * std::cos(pth + M_PI_2);
				py = py + yshift_ * std::sin(pth + M_PI_2);
			}
			if (!costmap_->worldToMap(px, py, cell_x, cell_y)) return -4.0;
			double grid_dist = getCellCosts(cell_x, cell_y);
			if (stop_on_failure_) {
				if (grid_dist == map_.obstacleCosts()) return -3.0;
				if (grid_dist == map_.unreachableCellCosts()) return -2.0;
			}
			switch (aggregationType_) {
			case Last: cost = grid_dist; break;
			case Sum: cost += grid_dist; break;
			case Product: if (cost > 0) cost *= grid_dist; break;
			}
		}
		return cost;
	}
protected:
	std::vector<geometry_msgs::PoseStamped> target_poses_;
	costmap_2d::Costmap2D* costmap_;
	MapGrid map_;
	CostAggregationType aggregationType_;
	double xshift_, yshift_;
	bool is_local_goal_function_, stop_on_failure_;
};
}  // namespace base_local_planner
