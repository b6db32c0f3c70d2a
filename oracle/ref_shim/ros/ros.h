// Stand-in for ros/ros.h: ros::Time / ros::Duration (only used for diagnostics timing on the path).
#pragma once
#include <chrono>
#include <cstdint>
#include <ros/console.h>
namespace ros {
class Duration {
public:
	explicit Duration(int64_t ns = 0) : ns_(ns) {}
	int64_t toNSec() const { return ns_; }
	double toSec() const { return 1e-9 * (double)ns_; }
private:
	int64_t ns_;
};
class Time {
public:
	Time() : ns_(0) {}
	static Time now() {
		Time t;
		t.ns_ = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
		return t;
	}
	Duration operator-(const Time& o) const { return Duration(ns_ - o.ns_); }
	double toSec() const { return 1e-9 * (double)ns_; }
private:
	int64_t ns_;
};
}  // namespace ros
