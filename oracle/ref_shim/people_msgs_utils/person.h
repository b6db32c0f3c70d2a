// Stand-in for people_msgs_utils/person.h: data getters only (the members the path reads).
#pragma once
#include <geometry_msgs/Pose.h>
#include <cmath>
#include <string>
namespace people_msgs_utils {
class Person {
public:
	Person(double x, double y, double yaw, double vx, double vy, double vth, double cxx, double cxy, double cyx, double cyy)
	    : x_(x), y_(y), yaw_(yaw), vx_(vx), vy_(vy), vth_(vth), cxx_(cxx), cxy_(cxy), cyx_(cyx), cyy_(cyy) {}
	geometry_msgs::Pose getPose() const {
		geometry_msgs::Pose p;
		p.position.x = x_;
		p.position.y = y_;
		p.orientation.z = std::sin(yaw_ / 2);
		p.orientation.w = std::cos(yaw_ / 2);
		return p;
	}
	double getPositionX() const { return x_; }
	double getPositionY() const { return y_; }
	double getVelocityX() const { return vx_; }
	double getVelocityY() const { return vy_; }
	double getVelocityTheta() const { return vth_; }
	double getCovariancePoseXX() const { return cxx_; }
	double getCovariancePoseXY() const { return cxy_; }
	double getCovariancePoseYX() const { return cyx_; }
	double getCovariancePoseYY() const { return cyy_; }
	double getReliability() const { return 1.0; }
	std::string getName() const { return "person"; }
protected:
	double x_, y_, yaw_, vx_, vy_, vth_, cxx_, cxy_, cyx_, cyy_;
};
}  // namespace people_msgs_utils
