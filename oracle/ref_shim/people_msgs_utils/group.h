// Stand-in for people_msgs_utils/group.h: data getters only.
#pragma once
#include <geometry_msgs/Pose.h>
#include <cmath>
#include <string>
namespace people_msgs_utils {
class Group {
public:
	Group(double x, double y, double yaw, double span_x, double span_y, double cxx, double cxy, double cyy)
	    : x_(x), y_(y), yaw_(yaw), span_x_(span_x), span_y_(span_y), cxx_(cxx), cxy_(cxy), cyy_(cyy) {}
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
	double getSpanX() const { return span_x_; }
	double getSpanY() const { return span_y_; }
	double getCovariancePoseXX() const { return cxx_; }
	double getCovariancePoseXY() const { return cxy_; }
	double getCovariancePoseYX() const { return cxy_; }
	double getCovariancePoseYY() const { return cyy_; }
	std::string getName() const { return "group"; }
protected:
	double x_, y_, yaw_, span_x_, span_y_, cxx_, cxy_, cyy_;
};
}  // namespace people_msgs_utils
