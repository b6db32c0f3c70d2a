// Stand-in for the upstream base_local_planner::PreferForwardCostFunction (backward_costs_): published formula.
#pragma once
#include <base_local_planner/trajectory_cost_function.h>
#include <cmath>
namespace base_local_planner {
class PreferForwardCostFunction : public TrajectoryCostFunction {
public:
	PreferForwardCostFunction(double penalty) : penalty_(penalty) {}
	bool prepare() override { return true; }
	void setPenalty(double penalty) { penalty_ = penalty; }
	double scoreTrajectory(Trajectory& traj) override {
		if (traj.xv_ < 0.0) return penalty_;                                  // backward motions bad on a robot without backward sensors
		if (traj.xv_ < 0.1 && std::fabs(traj.thetav_) < 0.2) return penalty_;  // strafing / barely moving
		return std::fabs(traj.thetav_) * 10;                                   // the more we rotate, the less we progress forward
	}
private:
	double penalty_;
};
}  // namespace base_local_planner
