// GPU test of the C++ adapter: the reference's call pattern (generator + critic list + SimpleScoredSamplingPlanner)
// must return the same winner as a direct hmp_plan() call. Run by tests/test_adapter_gpu.py on the GPU box.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include <base_local_planner/simple_scored_sampling_planner.h>

#include "gpu_social_trajectory_generator.h"

using namespace humap_local_planner_b200;

static HmpParams makeParams(double res) {
	HmpParams p;
	std::memset(&p, 0, sizeof(p));
	p.limits = {1.5, 0.1, 1.5, -0.1, 0.0, 0.0, 2.0, 0.4, 2.5, 0.0, 3.2, 0.4, 0, 0};
	p.general = {3.5, 0.1, 0.1, 0.1, 0.1, 1, 0};
	p.sfm.fov = 2.0; p.sfm.mass = 14.5; p.sfm.internal_force_factor = 0.75; p.sfm.static_interaction_force_factor = 4.9;
	p.sfm.dynamic_interaction_force_factor = 10.0; p.sfm.min_force = 5; p.sfm.max_force = 300; p.sfm.speed_desired = 1.29;
	p.sfm.relaxation_time = 0.54; p.sfm.an = -2.092; p.sfm.bn = 2.013; p.sfm.cn = 3.2421; p.sfm.ap = 1.5375; p.sfm.bp = 0.9876;
	p.sfm.cp = 0.4568; p.sfm.aw = 40.39; p.sfm.bw = 0.22452;
	p.fis = {100.0, 8.0, 3.31613 / 2.0, 0, 0};
	const double scales[HMP_NUM_COSTS] = {0.05, 15 * res, 25.5 * res, 8.5 * res, 8.0 * res, 6.0, 0.08, 3.0, 10, 17, 20, 30, 7.5, 10};
	for (int k = 0; k < HMP_NUM_COSTS; ++k) p.costs.scale[k] = scales[k];
	p.costs.occdist_separation = 0.025; p.costs.occdist_separation_kernel = 1;
	for (int g = 0; g < 4; ++g) { p.costs.xshift[g] = g >= 2 ? 0.325 : 0.0; p.costs.neighbour_kernel_size[g] = g >= 2 ? 3 : 0; p.costs.neighbour_cost_multiplier[g] = 3.0; }
	p.costs.unsat_max_trans_vel = 1.5; p.costs.unsat_max_vel_x = 1.5; p.costs.backward_penalty = 25; p.costs.ttc_collision_distance = 0.05;
	p.costs.hd_fov_person = 3.31613; p.costs.hd_person_model_radius = 0.4; p.costs.hd_robot_circumradius = 0.275; p.costs.hd_max_speed = 1.5;
	p.costs.ps_max_speed = 1.5; p.costs.ps_min_dist = 0.275;
	p.costs.hd_whole_horizon = p.costs.psi_whole_horizon = p.costs.fsi_whole_horizon = p.costs.ps_whole_horizon = 1;
	return p;
}

int main() {
	const int N = 200;
	const double res = 0.05, ox = -5.0, oy = -5.0;
	std::vector<uint8_t> cells((size_t)N * N, 0);
	for (int y = 120; y < 124; ++y) for (int x = 130; x < 140; ++x) cells[(size_t)y * N + x] = 254;   // a small wall, off the corridor
	// MapGrids: Manhattan distance to the plan cell row / to the goal cell (a wave front on an empty map)
	std::vector<double> path((size_t)N * N), goal((size_t)N * N), front((size_t)N * N);
	for (int y = 0; y < N; ++y) for (int x = 0; x < N; ++x) {
		int dxp = x < 100 ? 100 - x : (x > 180 ? x - 180 : 0);
		path[(size_t)y * N + x] = std::abs(y - 100) + dxp;
		goal[(size_t)y * N + x] = std::abs(y - 100) + std::abs(x - 180);
		front[(size_t)y * N + x] = std::abs(y - 100) + std::abs(x - 106);
	}
	std::vector<double> footprint;
	for (int i = 0; i < 16; ++i) { footprint.push_back(0.275 * std::cos(i * M_PI / 8)); footprint.push_back(0.275 * std::sin(i * M_PI / 8)); }

	std::vector<HmpObstacle> obs(3);
	std::memset(obs.data(), 0, sizeof(HmpObstacle) * obs.size());
	obs[0] = {0.19, 0.19, 0, 1.75, 1.1, 0, 0, 0, 0, 0, 0};
	obs[1] = {0.2, -0.19, 0, 2.0, -1.5, 0, 0, 0, 0, 0, 0};
	obs[2] = {0.27, 0.03, 0, 2.6, 0.3, 0, -0.6, 0.1, 0, 1, 0};   // a person walking towards the robot
	std::vector<HmpPerson> people(1);
	people[0] = {3.0, 0.35, std::atan2(0.1, -0.6), -0.6, 0.1, 0, 0.0025, 0, 0, 0.0025};
	HmpWorld w;
	std::memset(&w, 0, sizeof(w));
	w.vel_x = 0.3; w.goal_local_x = 4.0; w.goal_x = 8.0;
	w.obstacles = obs.data(); w.n_obstacles = 3; w.people = people.data(); w.n_people = 1;
	HmpSampling s;
	for (int a = 0; a < HMP_NUM_AMPLIFIERS; ++a) { s.amp_min[a] = s.amp_max[a] = 1.0; s.amp_granularity[a] = 1.0; }
	s.amp_min[HMP_AMP_AN] = -0.4; s.amp_max[HMP_AMP_AN] = 1.0; s.amp_granularity[HMP_AMP_AN] = 0.7;
	s.amp_min[HMP_AMP_AW] = 0.5; s.amp_max[HMP_AMP_AW] = 1.0; s.amp_granularity[HMP_AMP_AW] = 0.5;
	s.amp_min[HMP_AMP_BW] = 1.0; s.amp_max[HMP_AMP_BW] = 4.5; s.amp_granularity[HMP_AMP_BW] = 3.5;   // 3 x 2 x 2 = 12 candidates

	try {
		GpuSocialTrajectoryGenerator gen(0);
		HmpParams P = makeParams(res);
		gen.setParameters(P);
		gen.setCostmap(cells.data(), N, N, ox, oy, res);
		gen.setMapGrid(HMP_GRID_PATH, path.data(), 0);
		gen.setMapGrid(HMP_GRID_GOAL, goal.data(), 0);
		gen.setMapGrid(HMP_GRID_ALIGNMENT, path.data(), 0);
		gen.setMapGrid(HMP_GRID_GOAL_FRONT, front.data(), 0);
		gen.setFootprint(footprint);

		GpuPrecomputedCostFunction cost;
		std::vector<base_local_planner::TrajectoryCostFunction*> critics{&cost};
		std::vector<base_local_planner::TrajectorySampleGenerator*> gens{&gen};
		base_local_planner::SimpleScoredSamplingPlanner planner(gens, critics, -1, true);

		// 1) winner only (production mode)
		gen.initialise(w, s, false);
		base_local_planner::Trajectory result;
		result.cost_ = -7;   // humap_planner.cpp:1364
		std::vector<base_local_planner::Trajectory> explored;
		bool ok = planner.findBestTrajectory(result, &explored);
		const HmpResult& r = gen.result();
		std::printf("winner-only: ok=%d cost=%.9f xv=%.6f thv=%.6f points=%u explored=%zu | hmp best=%d total=%.9f n=%d gen=%d valid=%d\n",
		            ok, result.cost_, result.xv_, result.thetav_, result.getPointsSize(), explored.size(), r.best_index, r.best_total,
		            r.n_candidates, r.n_generated, r.n_valid);
		if (!ok || r.n_candidates != 12 || result.cost_ != r.best_total || result.getPointsSize() != 35 || explored.size() != 1 ||
		    result.xv_ != r.xv || result.thetav_ != r.thetav || std::fabs(result.time_delta_ - 0.1) > 1e-12) {
			std::printf("ADAPTER_TEST_FAIL winner-only\n");
			return 1;
		}
		// 2) explore-all (diagnostics mode): same winner, every generated candidate reported with its total
		gen.initialise(w, s, true);
		base_local_planner::Trajectory result2;
		result2.cost_ = -7;
		explored.clear();
		ok = planner.findBestTrajectory(result2, &explored);
		double min_cost = -1;
		for (auto& t : explored) if (t.cost_ >= 0 && (min_cost < 0 || t.cost_ < min_cost)) min_cost = t.cost_;
		std::printf("explore-all: ok=%d cost=%.9f explored=%zu min=%.9f\n", ok, result2.cost_, explored.size(), min_cost);
		if (!ok || result2.cost_ != result.cost_ || (int)explored.size() != gen.result().n_generated || min_cost != result.cost_ ||
		    result2.getPointsSize() != 35) {
			std::printf("ADAPTER_TEST_FAIL explore-all\n");
			return 1;
		}
		double x0, y0, t0, x1, y1, t1;
		result.getPoint(0, x0, y0, t0);
		result.getEndpoint(x1, y1, t1);
		if (x0 != 0.0 || y0 != 0.0 || !(x1 > 0.5)) {
			std::printf("ADAPTER_TEST_FAIL trajectory shape (%f %f -> %f %f)\n", x0, y0, x1, y1);
			return 1;
		}
		// 3) both generators of the reference's pool behind the one adapter generator
		HmpEquisampled eq = {1, 5, 1, 10, 0.1, 1, 0};
		gen.setEquisampled(&eq);
		gen.initialise(w, s, true);
		base_local_planner::Trajectory result3;
		result3.cost_ = -7;
		explored.clear();
		ok = planner.findBestTrajectory(result3, &explored);
		const HmpResult& r3 = gen.result();
		std::printf("pooled: ok=%d cost=%.9f best=%d n=%d social=%d explored=%zu\n", ok, result3.cost_, r3.best_index, r3.n_candidates,
		            r3.n_social, explored.size());
		if (!ok || r3.n_social != 12 || r3.n_candidates < 12 + 50 || result3.cost_ != r3.best_total || result3.cost_ > result.cost_ ||
		    (int)explored.size() != r3.n_generated || result3.getPointsSize() != 35) {
			std::printf("ADAPTER_TEST_FAIL pooled generators\n");
			return 1;
		}
		gen.setEquisampled(nullptr);
		std::printf("ADAPTER_TEST_OK\n");
	} catch (const std::exception& e) {
		std::printf("ADAPTER_TEST_FAIL exception: %s\n", e.what());
		return 1;
	}
	return 0;
}
