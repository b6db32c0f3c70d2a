// Minimal from-scratch stand-in for base_local_planner/trajectory.h (ROS navigation is not installable here).
// Only the members the sampling + scoring seam touches. In a ROS build this directory is
// NOT on the include path and the real headers are used.
#pragma once
#include <vector>

namespace base_local_planner {

class Trajectory {
public:
	Trajectory() : xv_(0), yv_(0), thetav_(0), cost_(-1), time_delta_(0) {}
	double xv_, yv_, thetav_;
	double cost_;
	double time_delta_;

	void getPoint(unsigned int index, double& x, double& y, double& th) const {
		x = x_pts_[index];
		y = y_pts_[index];
		th = th_pts_[index];
	}
	void setPoint(unsigned int index, double x, double y, double th) {
		x_pts_[index] = x;
		y_pts_[index] = y;
		th_pts_[index] = th;
	}
	void addPoint(double x, double y, double th) {
		x_pts_.push_back(x);
		y_pts_.push_back(y);
		th_pts_.push_back(th);
	}
	void getEndpoint(double& x, double& y, double& th) const {
		x = x_pts_.back();
		y = y_pts_.back();
		th = th_pts_.back();
	}
	void resetPoints() {
		x_pts_.clear();
		y_pts_.clear();
		th_pts_.clear();
	}
	unsigned int getPointsSize() const { return (unsigned int)x_pts_.size(); }

private:
	std::vector<double> x_pts_, y_pts_, th_pts_;
};

}  // namespace base_local_planner
