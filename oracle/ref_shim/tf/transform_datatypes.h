// Stand-in for tf/transform_datatypes.h: Vector3, Quaternion, Pose, Stamped<>, getYaw.
#pragma once
#include <cmath>
#include <string>
namespace tf {
class Vector3 {
public:
	Vector3() : v{0, 0, 0} {}
	Vector3(double x, double y, double z) : v{x, y, z} {}
	double getX() const { return v[0]; }
	double getY() const { return v[1]; }
	double getZ() const { return v[2]; }
	double x() const { return v[0]; }
	double y() const { return v[1]; }
	double z() const { return v[2]; }
private:
	double v[3];
};
typedef Vector3 Point;
class Quaternion {
public:
	Quaternion() : q{0, 0, 0, 1} {}
	Quaternion(double x, double y, double z, double w) : q{x, y, z, w} {}
	double x() const { return q[0]; }
	double y() const { return q[1]; }
	double z() const { return q[2]; }
	double w() const { return q[3]; }
private:
	double q[4];
};
class Transform {
public:
	const Vector3& getOrigin() const { return o; }
	Quaternion getRotation() const { return r; }
	void setOrigin(const Vector3& v) { o = v; }
	void setRotation(const Quaternion& q) { r = q; }
private:
	Vector3 o;
	Quaternion r;
};
typedef Transform Pose;
template <typename T>
class Stamped : public T {
public:
	double stamp_ = 0.0;
	std::string frame_id_;
};
// yaw of a (planar) rotation: atan2(2(wz + xy), 1 - 2(y^2 + z^2))
inline double getYaw(const Quaternion& q) {
	return std::atan2(2.0 * (q.w() * q.z() + q.x() * q.y()), 1.0 - 2.0 * (q.y() * q.y() + q.z() * q.z()));
}
inline Quaternion createQuaternionFromYaw(double yaw) { return Quaternion(0, 0, std::sin(yaw / 2), std::cos(yaw / 2)); }
}  // namespace tf
