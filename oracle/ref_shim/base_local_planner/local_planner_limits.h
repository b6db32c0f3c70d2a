// Stand-in for base_local_planner/local_planner_limits.h: the published plain-data limits record.
#pragma once
namespace base_local_planner {
class LocalPlannerLimits {
public:
	double max_vel_trans = 0, min_vel_trans = 0;
	double max_vel_x = 0, min_vel_x = 0;
	double max_vel_y = 0, min_vel_y = 0;
	double max_vel_theta = 0, min_vel_theta = 0;
	double acc_lim_x = 0, acc_lim_y = 0, acc_lim_theta = 0, acc_lim_trans = 0;
	bool prune_plan = false;
	double xy_goal_tolerance = 0, yaw_goal_tolerance = 0;
	double trans_stopped_vel = 0, theta_stopped_vel = 0;
	bool restore_defaults = false;
};
}  // namespace base_local_planner
