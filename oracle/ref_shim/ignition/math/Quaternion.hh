// Stand-in for ignition/math/Quaternion.hh (ignition-math 4), double only: Euler <-> quaternion.
#pragma once
#include <ignition/math/Helpers.hh>
#include <ignition/math/Vector3.hh>
namespace ignition {
namespace math {
template <typename T>
class Quaternion {
public:
	Quaternion() : qw(1), qx(0), qy(0), qz(0) {}
	Quaternion(const T& w, const T& x, const T& y, const T& z) : qw(w), qx(x), qy(y), qz(z) {}
	Quaternion(const T& roll, const T& pitch, const T& yaw) { Euler(roll, pitch, yaw); }
	void Normalize() {
		T s = std::sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
		if (equal<T>(s, static_cast<T>(0))) {
			qw = T(1);
			qx = qy = qz = T(0);
		} else {
			qw /= s;
			qx /= s;
			qy /= s;
			qz /= s;
		}
	}
	void Euler(T roll, T pitch, T yaw) {
		T phi = roll / T(2), the = pitch / T(2), psi = yaw / T(2);
		qw = std::cos(phi) * std::cos(the) * std::cos(psi) + std::sin(phi) * std::sin(the) * std::sin(psi);
		qx = std::sin(phi) * std::cos(the) * std::cos(psi) - std::cos(phi) * std::sin(the) * std::sin(psi);
		qy = std::cos(phi) * std::sin(the) * std::cos(psi) + std::sin(phi) * std::cos(the) * std::sin(psi);
		qz = std::cos(phi) * std::cos(the) * std::sin(psi) - std::sin(phi) * std::sin(the) * std::cos(psi);
		Normalize();
	}
	Vector3<T> Euler() const {
		Vector3<T> vec;
		const T tol = static_cast<T>(1e-15);
		Quaternion<T> c = *this;
		c.Normalize();
		T squ = c.qw * c.qw, sqx = c.qx * c.qx, sqy = c.qy * c.qy, sqz = c.qz * c.qz;
		T sarg = -2 * (c.qx * c.qz - c.qw * c.qy);
		if (sarg <= T(-1.0)) {
			vec.Y(T(-0.5 * IGN_PI));
		} else if (sarg >= T(1.0)) {
			vec.Y(T(0.5 * IGN_PI));
		} else {
			vec.Y(T(std::asin(sarg)));
		}
		if (std::abs(sarg - 1) < tol) {
			vec.Z(0);
			vec.X(T(std::atan2(2 * (c.qx * c.qy - c.qz * c.qw), squ - sqx + sqy - sqz)));
		} else if (std::abs(sarg + 1) < tol) {
			vec.Z(0);
			vec.X(T(std::atan2(-2 * (c.qx * c.qy - c.qz * c.qw), squ - sqx + sqy - sqz)));
		} else {
			vec.X(T(std::atan2(2 * (c.qy * c.qz + c.qw * c.qx), squ - sqx - sqy + sqz)));
			vec.Z(T(std::atan2(2 * (c.qx * c.qy + c.qw * c.qz), squ + sqx - sqy - sqz)));
		}
		return vec;
	}
	T Roll() const { return Euler().X(); }
	T Pitch() const { return Euler().Y(); }
	T Yaw() const { return Euler().Z(); }
	const T& W() const { return qw; }
	const T& X() const { return qx; }
	const T& Y() const { return qy; }
	const T& Z() const { return qz; }
	T& W() { return qw; }
	T& X() { return qx; }
	T& Y() { return qy; }
	T& Z() { return qz; }
	void W(T v) { qw = v; }
	void X(T v) { qx = v; }
	void Y(T v) { qy = v; }
	void Z(T v) { qz = v; }
private:
	T qw, qx, qy, qz;
};
typedef Quaternion<double> Quaterniond;
}  // namespace math
}  // namespace ignition
