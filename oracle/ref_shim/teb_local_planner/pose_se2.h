// Stand-in for teb_local_planner/pose_se2.h: x, y, theta carrier.
#pragma once
namespace teb_local_planner {
class PoseSE2 {
public:
	PoseSE2() : x_(0), y_(0), th_(0) {}
	PoseSE2(double x, double y, double theta) : x_(x), y_(y), th_(theta) {}
	double x() const { return x_; }
	double y() const { return y_; }
	double theta() const { return th_; }
private:
	double x_, y_, th_;
};
}  // namespace teb_local_planner
