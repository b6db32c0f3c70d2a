/*
 * gpu_social_trajectory_generator.h -- C++ adapter that keeps the base_local_planner plugin interfaces of the
 * sampling + scoring path while the work runs on the GPU behind the C ABI (include/hmp_planner.h).
 *
 *   GpuSocialTrajectoryGenerator : base_local_planner::TrajectorySampleGenerator
 *       replaces humap_local_planner::SocialTrajectoryGenerator (include/humap_local_planner/social_trajectory_generator.h:215-225)
 *   GpuPrecomputedCostFunction   : base_local_planner::TrajectoryCostFunction
 *       replaces the 14 critics registered at src/humap_planner.cpp:68-82: the GPU has already evaluated them in the
 *       reference order, so scoreTrajectory() returns the trajectory's precomputed weighted total (negative codes kept)
 *
 * Used exactly like the reference uses its generator + critics:
 *       gen.initialise(world, sampling);                                    // humap_planner.cpp:1307-1314
 *       SimpleScoredSamplingPlanner planner({&gen}, critics = {&cost}, -1, true);
 *       planner.findBestTrajectory(result_traj, &traj_explored);            // humap_planner.cpp:1367
 * The first hasMoreTrajectories() after initialise() triggers ONE hmp_plan() call (all candidates rolled out, scored and
 * the argmin selected on the device); the generator then yields the winner first (so that it stays the strict minimum in
 * the scored-sampling loop) and, when explore_all is on, every other generated candidate with its total for traj_explored_.
 */
#pragma once

#include <string>
#include <vector>

#include <base_local_planner/trajectory.h>
#include <base_local_planner/trajectory_cost_function.h>
#include <base_local_planner/trajectory_sample_generator.h>

#include "hmp_planner.h"

namespace humap_local_planner_b200 {

class GpuSocialTrajectoryGenerator : public base_local_planner::TrajectorySampleGenerator {
public:
	explicit GpuSocialTrajectoryGenerator(int device_id = 0);
	~GpuSocialTrajectoryGenerator() override;
	GpuSocialTrajectoryGenerator(const GpuSocialTrajectoryGenerator&) = delete;
	GpuSocialTrajectoryGenerator& operator=(const GpuSocialTrajectoryGenerator&) = delete;

	/// HumapPlanner::reconfigure / updateCostParameters / updateLocalCosts: throws std::runtime_error on failure
	void setParameters(const HmpParams& params);
	void setCostmap(const uint8_t* cells, int size_x, int size_y, double origin_x, double origin_y, double resolution);
	void setMapGrid(int grid, const double* target_dist, double highest_valid_cost_prev);
	void setFootprint(const std::vector<double>& xy);
	/// generator_vel_space_ (humap_planner.cpp:196-203, :1317-1361): when enabled, the equisampled-velocity candidates are
	/// rolled out and scored on the device in the same pool, after the social ones (generator_list order, :85-88), so this
	/// one generator stands for both entries of the reference's generator list; nullptr turns them off
	void setEquisampled(const HmpEquisampled* eq);

	/// SocialTrajectoryGenerator::initialise: stores the cycle's world + sampling; the plan runs lazily
	void initialise(const HmpWorld& world, const HmpSampling& sampling, bool explore_all = false);

	bool hasMoreTrajectories() override;
	bool nextTrajectory(base_local_planner::Trajectory& traj) override;

	/// result of the cycle (valid after the first hasMoreTrajectories())
	const HmpResult& result() const { return result_; }
	const std::string& lastError() const { return error_; }

private:
	void plan();
	HmpContext* ctx_;
	HmpWorld world_{};
	HmpSampling sampling_{};
	std::vector<HmpObstacle> obstacles_;
	std::vector<HmpPerson> people_;
	std::vector<HmpGroup> groups_;
	bool explore_all_ = false;
	bool planned_ = false;
	bool failed_ = false;
	HmpResult result_{};
	std::vector<double> best_poses_;
	std::vector<double> totals_;
	std::vector<int32_t> order_;       // candidate indices to yield, winner first
	size_t next_ = 0;
	// chunk of explained candidates (explore_all)
	std::vector<double> chunk_poses_, chunk_seeds_;
	std::vector<int32_t> chunk_nposes_;
	size_t chunk_begin_ = 0, chunk_end_ = 0;
	std::string error_;
};

class GpuPrecomputedCostFunction : public base_local_planner::TrajectoryCostFunction {
public:
	GpuPrecomputedCostFunction() : base_local_planner::TrajectoryCostFunction(1.0) {}
	bool prepare() override { return true; }
	/// the generator stored the device's weighted total in traj.cost_
	double scoreTrajectory(base_local_planner::Trajectory& traj) override { return traj.cost_; }
};

}  // namespace humap_local_planner_b200
