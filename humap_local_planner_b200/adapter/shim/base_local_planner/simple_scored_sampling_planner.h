// Stand-in for base_local_planner::SimpleScoredSamplingPlanner with the 4-argument constructor of the author's
// fork ("always use all generators", reference src/humap_planner.cpp:90-95). Semantics per SURVEY.md Appendix B.
#pragma once
#include <vector>

#include <base_local_planner/trajectory.h>
#include <base_local_planner/trajectory_cost_function.h>
#include <base_local_planner/trajectory_sample_generator.h>

namespace base_local_planner {

class SimpleScoredSamplingPlanner {
public:
	SimpleScoredSamplingPlanner() : max_samples_(-1), use_all_generators_(true) {}
	SimpleScoredSamplingPlanner(std::vector<TrajectorySampleGenerator*> gen_list, std::vector<TrajectoryCostFunction*>& critics,
	                            int max_samples = -1, bool use_all_generators = false)
	    : gen_list_(gen_list), critics_(critics), max_samples_(max_samples), use_all_generators_(use_all_generators) {}

	double scoreTrajectory(Trajectory& traj, double best_traj_cost) {
		double traj_cost = 0;
		for (TrajectoryCostFunction* score_function_p : critics_) {
			if (score_function_p->getScale() == 0) continue;
			double cost = score_function_p->scoreTrajectory(traj);
			if (cost < 0) {
				traj_cost = cost;
				break;
			}
			if (cost != 0) cost *= score_function_p->getScale();
			traj_cost += cost;
			if (best_traj_cost > 0) {
				if (traj_cost > best_traj_cost) break;
			}
		}
		return traj_cost;
	}

	bool findBestTrajectory(Trajectory& traj, std::vector<Trajectory>* all_explored = 0) {
		Trajectory loop_traj;
		Trajectory best_traj;
		double loop_traj_cost, best_traj_cost = -1;
		bool gen_success;
		int count;
		for (TrajectoryCostFunction* loop_critic_p : critics_) {
			if (loop_critic_p->prepare() == false) return false;
		}
		for (TrajectorySampleGenerator* gen_ : gen_list_) {
			count = 0;
			while (gen_->hasMoreTrajectories()) {
				gen_success = gen_->nextTrajectory(loop_traj);
				if (gen_success == false) continue;
				loop_traj_cost = scoreTrajectory(loop_traj, best_traj_cost);
				if (all_explored != NULL) {
					loop_traj.cost_ = loop_traj_cost;
					all_explored->push_back(loop_traj);
				}
				if (loop_traj_cost >= 0) {
					if (best_traj_cost < 0 || loop_traj_cost < best_traj_cost) {
						best_traj_cost = loop_traj_cost;
						best_traj = loop_traj;
					}
				}
				count++;
				if (max_samples_ > 0 && count >= max_samples_) break;
			}
			if (best_traj_cost >= 0) {
				traj.xv_ = best_traj.xv_;
				traj.yv_ = best_traj.yv_;
				traj.thetav_ = best_traj.thetav_;
				traj.cost_ = best_traj_cost;
				traj.time_delta_ = best_traj.time_delta_;
				traj.resetPoints();
				double px, py, pth;
				for (unsigned int i = 0; i < best_traj.getPointsSize(); i++) {
					best_traj.getPoint(i, px, py, pth);
					traj.addPoint(px, py, pth);
				}
			}
			if (best_traj_cost >= 0 && !use_all_generators_) break;
		}
		return best_traj_cost >= 0;
	}

private:
	std::vector<TrajectorySampleGenerator*> gen_list_;
	std::vector<TrajectoryCostFunction*> critics_;
	int max_samples_;
	bool use_all_generators_;
};

}  // namespace base_local_planner
