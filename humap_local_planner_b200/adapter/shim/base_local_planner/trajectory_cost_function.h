#pragma once
#include <base_local_planner/trajectory.h>

namespace base_local_planner {

class TrajectoryCostFunction {
public:
	virtual bool prepare() = 0;
	virtual double scoreTrajectory(Trajectory& traj) = 0;
	double getScale() { return scale_; }
	void setScale(double scale) { scale_ = scale; }
	virtual ~TrajectoryCostFunction() {}

protected:
	TrajectoryCostFunction(double scale = 1.0) : scale_(scale) {}

private:
	double scale_;
};

}  // namespace base_local_planner
