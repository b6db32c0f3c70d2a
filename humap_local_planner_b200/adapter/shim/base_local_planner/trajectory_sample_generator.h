#pragma once
#include <base_local_planner/trajectory.h>

namespace base_local_planner {

class TrajectorySampleGenerator {
public:
	virtual bool hasMoreTrajectories() = 0;
	virtual bool nextTrajectory(Trajectory& traj) = 0;
	virtual ~TrajectorySampleGenerator() {}

protected:
	TrajectorySampleGenerator() {}
};

}  // namespace base_local_planner
