/*
 * hmp_oracle.h -- C interface of the CPU oracle (see hmp_oracle.cpp). TEST INFRASTRUCTURE ONLY:
 * nothing under humap_local_planner_b200/ may include, link or load this.
 */
#ifndef HMP_ORACLE_H_
#define HMP_ORACLE_H_

#include "../include/hmp_planner.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcPlanInput {
	const HmpParams* params;
	const HmpWorld* world;
	const HmpSampling* sampling;
	const HmpSample* extra;
	int32_t n_extra;
	/* costmap */
	int32_t size_x, size_y;
	const uint8_t* cells;
	double origin_x, origin_y, resolution;
	/* MapGridCostFunction::prepare() outputs */
	const double* target_dist[HMP_NUM_MAPGRIDS];
	double highest_valid_cost_prev[HMP_NUM_MAPGRIDS];
	/* footprint */
	const double* footprint_xy;
	int32_t n_footprint;
	/* 1: reference semantics (SimpleScoredSamplingPlanner stops summing once worse than the best);
	 * 0: full totals for every candidate (what the GPU path reports in hmp_get_explored_totals) */
	int32_t early_exit;
	/* candidate sub-range [cand_begin, cand_end) to evaluate; cand_end <= 0 means all (used by the
	 * multi-threaded CPU baseline) */
	int32_t cand_begin, cand_end;
	/* second generator of the pool (may be NULL): equisampled velocities, appended after the social candidates */
	const HmpEquisampled* equisampled;
} OrcPlanInput;

typedef struct OrcPlanOutput {
	HmpResult result;
	/* all optional (NULL to skip); C = number of candidates, T = number of steps */
	double* totals;      /* [C] weighted total or negative code (-1 = generator rejected)       */
	double* costs;       /* [C][HMP_NUM_COSTS] raw critic outputs, NaN = not evaluated           */
	double* seeds;       /* [C][3]                                                               */
	double* poses;       /* [C][T][3]                                                            */
	int32_t* n_poses;    /* [C]                                                                  */
	int32_t* generated;  /* [C]                                                                  */
	double* best_poses;  /* [T][3]                                                               */
	double* forces;      /* [T][8] per-step forces of candidate `forces_candidate`               */
	int32_t forces_candidate;
	int32_t _pad;
} OrcPlanOutput;

int orc_plan(const OrcPlanInput* in, OrcPlanOutput* out);
int orc_score_trajectory(const OrcPlanInput* in, const double* poses, int n, const double seed[3], double* raw_costs,
                         double* total, double* hv_out);
/* HumapPlanner::computeCellCost for every cell (src/humap_planner.cpp:535-576): cloud6[cy * size_x + cx] = {total, path,
 * goal, layered, alignment, goal_front} floats, valid = 0 where the reference returns false */
int orc_cost_cloud(const OrcPlanInput* in, float* cloud6, uint8_t* valid);
/* HumapPlanner::createEnvironmentModel (src/humap_planner.cpp:930-1052); same contract as hmp_build_environment */
int orc_build_environment(const HmpEnvParams* env, const double robot_pose[3], const double pose_ref[3], const HmpShape* shapes,
                          int32_t n_shapes, const double* vertices_xy, int32_t n_vertices, const HmpPerson* people, int32_t n_people,
                          const HmpGroup* groups, int32_t n_groups, HmpObstacle* obstacles_out, int32_t* n_obstacles_out,
                          int32_t* people_selected, int32_t* n_people_selected, int32_t* groups_selected, int32_t* n_groups_selected);
/* HumapPlanner::computeForceAtPosition (src/humap_planner.cpp:652-678) per position; same contract as hmp_compute_force_grid */
int orc_force_grid(const HmpParams* P, const HmpEnvParams* env, const HmpWorld* world, const double* positions_xy, int32_t n_positions,
                   const HmpShape* shapes, int32_t n_shapes, const double* vertices_xy, int32_t n_vertices, double* forces_out);
int orc_num_candidates(const HmpSampling* sampling, int n_extra);
/* velocity samples of the equisampled generator for this cycle: out[n][3] floats stored as doubles; returns n */
int orc_equisampled_samples(const HmpParams* P, const HmpWorld* w, const HmpEquisampled* eq, double* out, int cap);
int orc_num_steps(const HmpParams* P, const HmpWorld* w);
int orc_samples(const HmpSampling* sampling, const HmpSample* extra, int n_extra, HmpSample* out);
void orc_mapgrid_compute(const uint8_t* cells, int size_x, int size_y, double origin_x, double origin_y,
                         double resolution, const double* plan_xy, int n_plan, int local_goal, double* target_dist);

#ifdef __cplusplus
}
#endif
#endif
