// Stand-in for ignition/math/Pose3.hh (ignition-math 4): position + orientation carrier.
#pragma once
#include <ignition/math/Quaternion.hh>
#include <ignition/math/Vector3.hh>
namespace ignition {
namespace math {
template <typename T>
class Pose3 {
public:
	Pose3() {}
	Pose3(const Vector3<T>& pos, const Quaternion<T>& rot) : p(pos), q(rot) {}
	Pose3(T x, T y, T z, T roll, T pitch, T yaw) : p(x, y, z), q(roll, pitch, yaw) {}
	Pose3(T x, T y, T z, T qw, T qx, T qy, T qz) : p(x, y, z), q(qw, qx, qy, qz) {}
	const Vector3<T>& Pos() const { return p; }
	Vector3<T>& Pos() { return p; }
	const Quaternion<T>& Rot() const { return q; }
	Quaternion<T>& Rot() { return q; }
private:
	Vector3<T> p;
	Quaternion<T> q;
};
typedef Pose3<double> Pose3d;
}  // namespace math
}  // namespace ignition
