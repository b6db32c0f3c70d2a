/*
 * from_reference_test.cpp -- TEST INFRASTRUCTURE for humap_local_planner_b200/adapter/from_reference.h.
 * Built into tests/_build/libfrom_reference_test.so against the reference's own headers (+ oracle/ref_shim stand-ins) and
 * linked with oracle/_ref/libhmp_ref.so (the reference's geometry / World objects) and the adapter + libhmp_planner.so.
 *
 *   fr_roundtrip   flat inputs -> the reference's own objects (HumapConfig, World via World::addObstacle in call order,
 *                  Person / Group, TrajectorySamplingParams) -> from_reference.h converters -> flat outputs. The CPU test
 *                  requires the outputs to equal the inputs (objects in the World's dynamic-then-static order).
 *   fr_plan        the reference's call pattern on the reference's types: setParameters(cfg ptrs...), initialise(world_model,
 *                  vel, sampling, limits, mass, true), SimpleScoredSamplingPlanner::findBestTrajectory -- GPU test: equals a
 *                  direct hmp_plan on the flat inputs.
 */
#include <cstring>

#include <base_local_planner/simple_scored_sampling_planner.h>

#include "../humap_local_planner_b200/adapter/from_reference.h"

namespace hlp = humap_local_planner;
using namespace humap_local_planner_b200;
using hlp::geometry::Pose;
using hlp::geometry::Vector;

namespace {

struct OpenConfig : hlp::HumapConfig {   // HumapConfigROS fills the protected members the same way (humap_config_ros.cpp)
	using hlp::HumapConfig::costs_;
	using hlp::HumapConfig::fis_;
	using hlp::HumapConfig::general_;
	using hlp::HumapConfig::limits_;
	using hlp::HumapConfig::sfm_;
	using hlp::HumapConfig::traj_gen_;
	using hlp::HumapConfig::traj_sampling_;
};

// inverse of toHmpParams: what a HumapConfig must hold so that the planner pushes `P` into its generator and critics
void fillConfig(OpenConfig& cfg, const HmpParams& P, const HmpSampling& S, double resolution) {
	auto& L = *cfg.limits_;
	L.max_vel_trans = P.limits.max_vel_trans; L.min_vel_trans = P.limits.min_vel_trans;
	L.max_vel_x = P.limits.max_vel_x; L.min_vel_x = P.limits.min_vel_x;
	L.max_vel_y = P.limits.max_vel_y; L.min_vel_y = P.limits.min_vel_y;
	L.max_vel_theta = P.limits.max_vel_theta; L.min_vel_theta = P.limits.min_vel_theta;
	L.acc_lim_x = P.limits.acc_lim_x; L.acc_lim_y = P.limits.acc_lim_y; L.acc_lim_theta = P.limits.acc_lim_theta;
	L.twist_rotation_compensation = P.limits.twist_rotation_compensation;
	L.maintain_vel_components_rate = P.limits.maintain_vel_components_rate != 0;
	auto& G = *cfg.general_;
	G.sim_time = P.general.sim_time; G.sim_granularity = P.general.sim_granularity;
	G.angular_sim_granularity = P.general.angular_sim_granularity; G.sim_period = P.general.sim_period;
	G.person_fov = P.costs.hd_fov_person / 2.0; G.person_model_radius = P.costs.hd_person_model_radius;
	auto& F = *cfg.sfm_;
	F.fov = P.sfm.fov; F.fov_factor_method = (unsigned)P.sfm.fov_factor_method; F.mass = P.sfm.mass;
	F.internal_force_factor = P.sfm.internal_force_factor; F.static_interaction_force_factor = P.sfm.static_interaction_force_factor;
	F.dynamic_interaction_force_factor = P.sfm.dynamic_interaction_force_factor; F.min_force = P.sfm.min_force; F.max_force = P.sfm.max_force;
	F.filter_forces = P.sfm.filter_forces != 0; F.disable_interaction_forces = P.sfm.disable_interaction_forces != 0;
	F.speed_desired = P.sfm.speed_desired; F.relaxation_time = P.sfm.relaxation_time;
	F.an = P.sfm.an; F.bn = P.sfm.bn; F.cn = P.sfm.cn; F.ap = P.sfm.ap; F.bp = P.sfm.bp; F.cp = P.sfm.cp; F.aw = P.sfm.aw; F.bw = P.sfm.bw;
	auto& I = *cfg.fis_;
	I.force_factor = P.fis.force_factor; I.human_action_range = P.fis.human_action_range; I.fov = P.fis.fov;
	I.fov_factor_method = (unsigned)P.fis.fov_factor_method;
	auto& C = *cfg.costs_;
	C.occdist_scale = P.costs.scale[HMP_COST_OBSTACLE];
	C.path_distance_scale = P.costs.scale[HMP_COST_PATH] / resolution;
	C.goal_distance_scale = P.costs.scale[HMP_COST_GOAL] / resolution;
	C.alignment_scale = P.costs.scale[HMP_COST_ALIGNMENT] / resolution;
	C.goal_front_scale = P.costs.scale[HMP_COST_GOAL_FRONT] / resolution;
	C.unsaturated_translation_scale = P.costs.scale[HMP_COST_UNSATURATED];
	C.backward_scale = P.costs.scale[HMP_COST_BACKWARD];
	C.ttc_scale = P.costs.scale[HMP_COST_TTC];
	C.heading_change_smoothness_scale = P.costs.scale[HMP_COST_HEADING_CHANGE];
	C.velocity_smoothness_scale = P.costs.scale[HMP_COST_VEL_SMOOTHNESS];
	C.heading_dir_scale = P.costs.scale[HMP_COST_HEADING_DIST];
	C.personal_space_scale = P.costs.scale[HMP_COST_PERSONAL_SPACE];
	C.fformation_space_scale = P.costs.scale[HMP_COST_FFORMATION];
	C.passing_speed_scale = P.costs.scale[HMP_COST_PASSING_SPEED];
	C.occdist_separation = P.costs.occdist_separation; C.occdist_separation_kernel = (unsigned short)P.costs.occdist_separation_kernel;
	C.occdist_sum_scores = P.costs.occdist_sum_scores != 0;
	C.forward_point_distance = P.costs.xshift[HMP_GRID_GOAL_FRONT];
	C.backward_penalty = P.costs.backward_penalty; C.ttc_rollout_time = P.costs.ttc_rollout_time;
	C.ttc_collision_distance = P.costs.ttc_collision_distance;
	C.unsaturated_translation_compute_whole_horizon = P.costs.unsat_whole_horizon != 0;
	C.heading_dir_compute_whole_horizon = P.costs.hd_whole_horizon != 0;
	C.personal_space_compute_whole_horizon = P.costs.psi_whole_horizon != 0;
	C.fformation_space_compute_whole_horizon = P.costs.fsi_whole_horizon != 0;
	C.passing_speed_compute_whole_horizon = P.costs.ps_whole_horizon != 0;
	auto& T = *cfg.traj_sampling_;
	double* f[HMP_NUM_AMPLIFIERS][3] = {
	    {&T.sfm_desired_speed_amplifier_min, &T.sfm_desired_speed_amplifier_max, &T.sfm_desired_speed_amplifier_granularity},
	    {&T.sfm_an_amplifier_min, &T.sfm_an_amplifier_max, &T.sfm_an_amplifier_granularity},
	    {&T.sfm_bn_amplifier_min, &T.sfm_bn_amplifier_max, &T.sfm_bn_amplifier_granularity},
	    {&T.sfm_cn_amplifier_min, &T.sfm_cn_amplifier_max, &T.sfm_cn_amplifier_granularity},
	    {&T.sfm_ap_amplifier_min, &T.sfm_ap_amplifier_max, &T.sfm_ap_amplifier_granularity},
	    {&T.sfm_bp_amplifier_min, &T.sfm_bp_amplifier_max, &T.sfm_bp_amplifier_granularity},
	    {&T.sfm_cp_amplifier_min, &T.sfm_cp_amplifier_max, &T.sfm_cp_amplifier_granularity},
	    {&T.sfm_aw_amplifier_min, &T.sfm_aw_amplifier_max, &T.sfm_aw_amplifier_granularity},
	    {&T.sfm_bw_amplifier_min, &T.sfm_bw_amplifier_max, &T.sfm_bw_amplifier_granularity},
	    {&T.fis_as_amplifier_min, &T.fis_as_amplifier_max, &T.fis_as_amplifier_granularity},
	};
	for (int a = 0; a < HMP_NUM_AMPLIFIERS; ++a) {
		*f[a][0] = S.amp_min[a];
		*f[a][1] = S.amp_max[a];
		*f[a][2] = S.amp_granularity[a];
	}
}

// HumapPlanner::plan (src/humap_planner.cpp:365-370) + createEnvironmentModel's World::addObstacle calls (:985-1036)
hlp::World buildWorld(const HmpWorld& hw) {
	Pose pose(hw.robot_x, hw.robot_y, hw.robot_yaw);
	Vector vel(hw.vel_x, hw.vel_y, hw.vel_th);
	Vector vel_glob;
	hlp::computeVelocityGlobal(vel, pose, vel_glob);
	hlp::World world(pose, vel_glob, Pose(hw.goal_local_x, hw.goal_local_y, hw.goal_local_yaw), Pose(hw.goal_x, hw.goal_y, hw.goal_yaw));
	for (int i = 0; i < hw.n_obstacles; ++i) {
		const HmpObstacle& o = hw.obstacles[i];
		world.addObstacle(Pose(o.robot_x, o.robot_y, o.robot_yaw), Pose(o.obj_x, o.obj_y, o.obj_yaw), Vector(o.vx, o.vy, o.vth), o.force_dynamic != 0);
	}
	return world;
}

void buildPeople(const HmpWorld& hw, const HmpParams& P, std::vector<hlp::Person>& people, std::vector<hlp::Group>& groups) {
	const unsigned int steps = (unsigned int)std::ceil(P.general.sim_time / P.general.sim_granularity);
	for (int i = 0; i < hw.n_people; ++i) {
		const HmpPerson& p = hw.people[i];
		people.push_back(hlp::Person(people_msgs_utils::Person(p.x, p.y, p.yaw, p.vx, p.vy, p.vth, p.cov_xx, p.cov_xy, p.cov_yx, p.cov_yy),
		                             P.general.sim_granularity, steps));
	}
	for (int i = 0; i < hw.n_groups; ++i) {
		const HmpGroup& g = hw.groups[i];
		groups.push_back(hlp::Group(people_msgs_utils::Group(g.x, g.y, g.yaw, g.span_x, g.span_y, g.cov_xx, g.cov_xy, g.cov_yy),
		                            P.general.sim_granularity, steps));
	}
}

}  // namespace

extern "C" {

// capacities: obstacles_out / people_out / groups_out hold at least the input counts
int fr_roundtrip(const HmpParams* P, const HmpWorld* hw, const HmpSampling* S, double resolution, double inscribed_radius, HmpParams* p_out,
                 HmpSampling* s_out, HmpWorld* w_out, HmpObstacle* obstacles_out, HmpPerson* people_out, HmpGroup* groups_out) {
	OpenConfig cfg;
	fillConfig(cfg, *P, *S, resolution);
	*p_out = toHmpParams(cfg, resolution, inscribed_radius);
	*s_out = toHmpSampling(*cfg.getTrajectorySampling());
	hlp::World world = buildWorld(*hw);
	std::vector<hlp::Person> people;
	std::vector<hlp::Group> groups;
	buildPeople(*hw, *P, people, groups);
	HmpWorldStorage st;
	toHmpWorld(world, Vector(hw->vel_x, hw->vel_y, hw->vel_th), people, groups, st);
	*w_out = st.world;
	std::memcpy(obstacles_out, st.obstacles.data(), st.obstacles.size() * sizeof(HmpObstacle));
	std::memcpy(people_out, st.people.data(), st.people.size() * sizeof(HmpPerson));
	std::memcpy(groups_out, st.groups.data(), st.groups.size() * sizeof(HmpGroup));
	w_out->obstacles = obstacles_out;
	w_out->people = people_out;
	w_out->groups = groups_out;
	return 0;
}

// The reference's call sequence on the reference's types, driven through the adapter. grids: 4 x size_x*size_y doubles.
int fr_plan(const HmpParams* P, const HmpWorld* hw, const HmpSampling* S, const uint8_t* cells, int size_x, int size_y, double ox, double oy,
            double resolution, const double* const grids[HMP_NUM_MAPGRIDS], const double hv_prev[HMP_NUM_MAPGRIDS], const double* footprint_xy,
            int n_footprint, double inscribed_radius, HmpResult* result_out, double* traj_cost_out, int* traj_points_out) {
	try {
		OpenConfig cfg;
		fillConfig(cfg, *P, *S, resolution);
		std::vector<hlp::Person> people_env_model;
		std::vector<hlp::Group> groups_env_model;
		GpuSocialTrajectoryGeneratorRef generator_social(people_env_model, groups_env_model, 0);
		GpuPrecomputedCostFunction gpu_costs;
		std::vector<base_local_planner::TrajectoryCostFunction*> critics{&gpu_costs};
		std::vector<base_local_planner::TrajectorySampleGenerator*> generator_list{&generator_social};
		base_local_planner::SimpleScoredSamplingPlanner scored_sampling_planner(generator_list, critics, -1, true);
		// reconfigure (:205-226) + updateCostParameters (:868-928)
		generator_social.setParameters(cfg.getSfm(), cfg.getFis(), cfg.getGeneral()->sim_time, cfg.getGeneral()->sim_granularity,
		                               cfg.getGeneral()->angular_sim_granularity, cfg.getGeneral()->sim_period,
		                               cfg.getLimits()->maintain_vel_components_rate, false, false, false, false);
		HmpParams params = toHmpParams(cfg, resolution, inscribed_radius);
		// updateLocalCosts (:1054-1141): the flat test input carries the cycle's x-shifts / scales already, reproduce them
		for (int g = 0; g < HMP_NUM_MAPGRIDS; ++g) params.costs.xshift[g] = P->costs.xshift[g];
		generator_social.setCostParameters(params);
		generator_social.setCostmap(cells, size_x, size_y, ox, oy, resolution);
		for (int g = 0; g < HMP_NUM_MAPGRIDS; ++g) generator_social.setMapGrid(g, grids[g], hv_prev[g]);
		generator_social.setFootprint(std::vector<double>(footprint_xy, footprint_xy + 2 * n_footprint));
		// plan (:365-370), createEnvironmentModel (:985-1036), planMovingRobot (:1305-1367)
		hlp::World world_model = buildWorld(*hw);
		buildPeople(*hw, *P, people_env_model, groups_env_model);
		Vector vel(hw->vel_x, hw->vel_y, hw->vel_th);
		generator_social.initialise(world_model, vel, *cfg.getTrajectorySampling(), cfg.getLimits(), cfg.getSfm()->mass, true);
		base_local_planner::Trajectory result_traj;
		result_traj.cost_ = -7;
		result_traj.resetPoints();
		std::vector<base_local_planner::Trajectory> traj_explored;
		scored_sampling_planner.findBestTrajectory(result_traj, &traj_explored);
		*result_out = generator_social.result();
		*traj_cost_out = result_traj.cost_;
		*traj_points_out = (int)result_traj.getPointsSize();
		return 0;
	} catch (const std::exception& e) {
		std::fprintf(stderr, "fr_plan: %s\n", e.what());
		return 1;
	}
}

}  // extern "C"
