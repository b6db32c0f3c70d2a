#include "gpu_social_trajectory_generator.h"

#include <algorithm>
#include <numeric>
#include <stdexcept>

namespace humap_local_planner_b200 {

namespace {
void check(int rc, const char* what) {
	if (rc != HMP_OK) throw std::runtime_error(std::string(what) + ": " + hmp_last_error());
}
constexpr size_t EXPLAIN_CHUNK = 4096;
}  // namespace

GpuSocialTrajectoryGenerator::GpuSocialTrajectoryGenerator(int device_id) : ctx_(hmp_create(device_id)) {
	// no CPU fallback: the planner must not start silently on a slower path
	if (!ctx_) throw std::runtime_error(std::string("hmp_create: ") + hmp_last_error());
}

GpuSocialTrajectoryGenerator::~GpuSocialTrajectoryGenerator() { hmp_destroy(ctx_); }

void GpuSocialTrajectoryGenerator::setParameters(const HmpParams& params) { check(hmp_set_params(ctx_, &params), "hmp_set_params"); }

void GpuSocialTrajectoryGenerator::setCostmap(const uint8_t* cells, int size_x, int size_y, double origin_x, double origin_y,
                                              double resolution) {
	check(hmp_set_costmap(ctx_, cells, size_x, size_y, origin_x, origin_y, resolution), "hmp_set_costmap");
}

void GpuSocialTrajectoryGenerator::setMapGrid(int grid, const double* target_dist, double highest_valid_cost_prev) {
	check(hmp_set_mapgrid(ctx_, grid, target_dist, highest_valid_cost_prev), "hmp_set_mapgrid");
}

void GpuSocialTrajectoryGenerator::setFootprint(const std::vector<double>& xy) {
	check(hmp_set_footprint(ctx_, xy.data(), (int32_t)(xy.size() / 2)), "hmp_set_footprint");
}

void GpuSocialTrajectoryGenerator::setEquisampled(const HmpEquisampled* eq) {
	if (hmp_set_equisampled(ctx_, eq) != HMP_OK) throw std::runtime_error(std::string("hmp_set_equisampled: ") + hmp_last_error());
}

void GpuSocialTrajectoryGenerator::initialise(const HmpWorld& world, const HmpSampling& sampling, bool explore_all) {
	// deep copy: the reference's generator copies the World too (social_trajectory_generator.cpp:89)
	obstacles_.assign(world.obstacles, world.obstacles + world.n_obstacles);
	people_.assign(world.people, world.people + world.n_people);
	groups_.assign(world.groups, world.groups + world.n_groups);
	world_ = world;
	world_.obstacles = obstacles_.data();
	world_.people = people_.data();
	world_.groups = groups_.data();
	sampling_ = sampling;
	explore_all_ = explore_all;
	planned_ = false;
	failed_ = false;
	next_ = 0;
	order_.clear();
	chunk_begin_ = chunk_end_ = 0;
}

void GpuSocialTrajectoryGenerator::plan() {
	planned_ = true;
	best_poses_.assign((size_t)HMP_MAX_STEPS * 3, 0.0);
	int rc = hmp_plan(ctx_, &world_, &sampling_, nullptr, 0, &result_, best_poses_.data(), HMP_MAX_STEPS);
	if (rc != HMP_OK) {
		error_ = hmp_last_error();
		failed_ = true;   // behaves like a generator without samples: findBestTrajectory() returns false, cost_ stays -7
		return;
	}
	if (result_.best_index >= 0) order_.push_back(result_.best_index);
	if (explore_all_) {
		totals_.assign((size_t)result_.n_candidates, 0.0);
		check(hmp_get_explored_totals(ctx_, totals_.data(), result_.n_candidates), "hmp_get_explored_totals");
		for (int32_t c = 0; c < result_.n_candidates; ++c) {
			if (c != result_.best_index && totals_[c] != -1.0) order_.push_back(c);   // -1: nextTrajectory() returned false
		}
	}
}

bool GpuSocialTrajectoryGenerator::hasMoreTrajectories() {
	if (!planned_) plan();
	return !failed_ && next_ < order_.size();
}

bool GpuSocialTrajectoryGenerator::nextTrajectory(base_local_planner::Trajectory& traj) {
	if (!hasMoreTrajectories()) return false;
	const size_t k = next_++;
	const int32_t cand = order_[k];
	traj.resetPoints();
	traj.time_delta_ = result_.time_delta;
	if (k == 0 && cand == result_.best_index) {
		traj.xv_ = result_.xv;
		traj.yv_ = result_.yv;
		traj.thetav_ = result_.thetav;
		traj.cost_ = result_.best_total;
		for (int i = 0; i < result_.n_poses; ++i) traj.addPoint(best_poses_[3 * i], best_poses_[3 * i + 1], best_poses_[3 * i + 2]);
		return true;
	}
	// explore_all: fetch the other candidates' poses in chunks through hmp_explain
	if (k < chunk_begin_ || k >= chunk_end_) {
		chunk_begin_ = k;
		chunk_end_ = std::min(order_.size(), k + EXPLAIN_CHUNK);
		const size_t n = chunk_end_ - chunk_begin_;
		const int T = hmp_num_steps(ctx_);
		chunk_poses_.assign(n * (size_t)T * 3, 0.0);
		chunk_seeds_.assign(n * 3, 0.0);
		chunk_nposes_.assign(n, 0);
		check(hmp_explain(ctx_, order_.data() + chunk_begin_, (int32_t)n, nullptr, chunk_seeds_.data(), chunk_poses_.data(),
		                  chunk_nposes_.data()), "hmp_explain");
	}
	const size_t j = k - chunk_begin_;
	const int T = hmp_num_steps(ctx_);
	traj.xv_ = chunk_seeds_[3 * j];
	traj.yv_ = chunk_seeds_[3 * j + 1];
	traj.thetav_ = chunk_seeds_[3 * j + 2];
	traj.cost_ = totals_[cand];
	for (int i = 0; i < chunk_nposes_[j]; ++i) {
		const double* p = &chunk_poses_[(j * (size_t)T + i) * 3];
		traj.addPoint(p[0], p[1], p[2]);
	}
	return true;
}

}  // namespace humap_local_planner_b200
