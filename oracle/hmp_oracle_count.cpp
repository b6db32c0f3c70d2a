/*
 * hmp_oracle_count.cpp -- the CPU oracle compiled a second time with `double` replaced by a counting scalar
 * (counted_double.h): the INSTRUMENTED floating-point operation count of the path per candidate, split into the rollout
 * (SocialTrajectoryGenerator::generateTrajectory, src/social_trajectory_generator.cpp:292-462) and the scoring
 * (SimpleScoredSamplingPlanner::scoreTrajectory over the 14 critics, src/humap_planner.cpp:868-928).
 *
 * TEST / MEASUREMENT INFRASTRUCTURE like the rest of oracle/: used by tools/count_flops.py (which writes
 * profiles/ r02_flop_count.json) and tests/test_flop_count.py. bench.py only READS the committed JSON.
 * The arithmetic is the oracle's own source text (hmp_oracle.cpp is included below, nothing is restated here), so the
 * counted build returns bit-identical totals -- orc_count_plan hands them back and the test compares them with orc_plan's.
 */
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <queue>
#include <string>
#include <vector>

#include "hmp_oracle.h"
#include "counted_double.h"

typedef double orc_f64;
#define HMP_ORACLE_COUNT 1
#define double hmp_count::cd
#include "hmp_oracle.cpp"
#undef double

extern "C" {

/* ops_rollout / ops_scoring: [hmp_count::N_OPS] sums over the evaluated candidates (add, mul, div, sqrt, exp, trig, atan2,
 * cmp, rnd); steps_rolled = candidate-steps the generator executed (a rejected candidate stops early), n_generated =
 * candidates that reached the critics; totals_out (optional) [n_idx] = weighted total or negative code, as orc_plan's. */
int orc_count_plan(const OrcPlanInput* in, const int32_t* cand_idx, int32_t n_idx, uint64_t* ops_rollout, uint64_t* ops_scoring,
                   uint64_t* steps_rolled, int32_t* n_generated, double* totals_out) {
	using hmp_count::cd;
	PlanState st;
	st.P = in->params;
	st.cm = Costmap{in->cells, in->size_x, in->size_y, cd(in->origin_x), cd(in->origin_y), cd(in->resolution)};
	for (int g = 0; g < HMP_NUM_MAPGRIDS; ++g) {
		MapGridCritic& m = st.grids[g];
		m.target_dist = reinterpret_cast<const cd*>(in->target_dist[g]);   // cd is a struct of one double
		m.size_x = in->size_x;
		m.size_y = in->size_y;
		m.xshift = in->params->costs.xshift[g];
		m.yshift = in->params->costs.yshift[g];
		m.stop_on_failure = in->params->costs.stop_on_failure[g] != 0;
		m.n_kernel_size = in->params->costs.neighbour_kernel_size[g];
		m.n_cost_multiplier = in->params->costs.neighbour_cost_multiplier[g];
		m.highest_valid_cost_prev = in->highest_valid_cost_prev[g];
		m.highest_valid_cost = 0.0;
	}
	st.footprint.assign(in->footprint_xy, in->footprint_xy + 2 * in->n_footprint);
	buildScene(st, *in->params, *in->world);
	auto samples = buildSamples(*in->sampling, in->extra, in->n_extra);
	FisEngine fis;
	const int T = computeStepsNumber(in->params->general, std::hypot(in->world->vel_x, in->world->vel_y), in->world->vel_th);
	hmp_count::Counters& c = hmp_count::counters();
	for (int k = 0; k < hmp_count::N_OPS; ++k) ops_rollout[k] = ops_scoring[k] = 0;
	*steps_rolled = 0;
	*n_generated = 0;
	BlpTrajectory traj;
	for (int i = 0; i < n_idx; ++i) {
		const int ci = cand_idx[i];
		if (ci < 0 || ci >= (int)samples.size()) return -1;
		hmp_count::Counters a = c;
		bool ok = generateTrajectory(*in->params, fis, st.world, st.vel_local, samples[ci], traj, nullptr);
		hmp_count::Counters b = c;
		for (int k = 0; k < hmp_count::N_OPS; ++k) ops_rollout[k] += b.n[k] - a.n[k];
		// a rejected rollout stops in the step whose twist violates the limits: traj.size() completed steps + that one
		*steps_rolled += ok ? (uint64_t)T : (uint64_t)traj.size() + 1;
		if (!ok) {
			if (totals_out) totals_out[i] = -1.0;
			continue;
		}
		(*n_generated)++;
		cd raw[HMP_NUM_COSTS];
		cd cost = scoreTrajectoryAll(st, traj, cd(-1.0), false, raw);
		hmp_count::Counters e = c;
		for (int k = 0; k < hmp_count::N_OPS; ++k) ops_scoring[k] += e.n[k] - b.n[k];
		if (totals_out) totals_out[i] = cost.v;
	}
	return 0;
}

int orc_count_num_ops(void) { return hmp_count::N_OPS; }

}  // extern "C"
