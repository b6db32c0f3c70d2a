// Stand-in for ignition/math/Vector3.hh (ignition-math 4), double only.
#pragma once
#include <ignition/math/Helpers.hh>
namespace ignition {
namespace math {
template <typename T>
class Vector3 {
public:
	Vector3() : d{0, 0, 0} {}
	Vector3(const T& x, const T& y, const T& z) : d{x, y, z} {}
	T X() const { return d[0]; }
	T Y() const { return d[1]; }
	T Z() const { return d[2]; }
	T& X() { return d[0]; }
	T& Y() { return d[1]; }
	T& Z() { return d[2]; }
	void X(const T& v) { d[0] = v; }
	void Y(const T& v) { d[1] = v; }
	void Z(const T& v) { d[2] = v; }
	void Set(T x = 0, T y = 0, T z = 0) { d[0] = x; d[1] = y; d[2] = z; }
	T Length() const { return std::sqrt(SquaredLength()); }
	T SquaredLength() const { return d[0] * d[0] + d[1] * d[1] + d[2] * d[2]; }
	// divides by the length unless it equals 0 within 1e-6
	Vector3 Normalize() {
		T l = Length();
		if (!equal<T>(l, static_cast<T>(0))) {
			d[0] /= l;
			d[1] /= l;
			d[2] /= l;
		}
		return *this;
	}
	Vector3 Normalized() const {
		Vector3 r = *this;
		r.Normalize();
		return r;
	}
	T Dot(const Vector3& v) const { return d[0] * v.d[0] + d[1] * v.d[1] + d[2] * v.d[2]; }
	Vector3 Cross(const Vector3& v) const {
		return Vector3(d[1] * v.d[2] - d[2] * v.d[1], d[2] * v.d[0] - d[0] * v.d[2], d[0] * v.d[1] - d[1] * v.d[0]);
	}
	T Distance(const Vector3& v) const { return (*this - v).Length(); }
	Vector3& operator=(T v) { d[0] = d[1] = d[2] = v; return *this; }
	Vector3 operator+(const Vector3& v) const { return Vector3(d[0] + v.d[0], d[1] + v.d[1], d[2] + v.d[2]); }
	const Vector3& operator+=(const Vector3& v) { d[0] += v.d[0]; d[1] += v.d[1]; d[2] += v.d[2]; return *this; }
	Vector3 operator+(const T s) const { return Vector3(d[0] + s, d[1] + s, d[2] + s); }
	friend Vector3 operator+(const T s, const Vector3& v) { return Vector3(v.d[0] + s, v.d[1] + s, v.d[2] + s); }
	const Vector3& operator+=(const T s) { d[0] += s; d[1] += s; d[2] += s; return *this; }
	Vector3 operator-() const { return Vector3(-d[0], -d[1], -d[2]); }
	Vector3 operator-(const Vector3& v) const { return Vector3(d[0] - v.d[0], d[1] - v.d[1], d[2] - v.d[2]); }
	const Vector3& operator-=(const Vector3& v) { d[0] -= v.d[0]; d[1] -= v.d[1]; d[2] -= v.d[2]; return *this; }
	Vector3 operator-(const T s) const { return Vector3(d[0] - s, d[1] - s, d[2] - s); }
	friend Vector3 operator-(const T s, const Vector3& v) { return Vector3(s - v.d[0], s - v.d[1], s - v.d[2]); }
	const Vector3& operator-=(const T s) { d[0] -= s; d[1] -= s; d[2] -= s; return *this; }
	const Vector3 operator/(const Vector3& v) const { return Vector3(d[0] / v.d[0], d[1] / v.d[1], d[2] / v.d[2]); }
	const Vector3& operator/=(const Vector3& v) { d[0] /= v.d[0]; d[1] /= v.d[1]; d[2] /= v.d[2]; return *this; }
	const Vector3 operator/(T s) const { return Vector3(d[0] / s, d[1] / s, d[2] / s); }
	const Vector3& operator/=(T s) { d[0] /= s; d[1] /= s; d[2] /= s; return *this; }
	Vector3 operator*(const Vector3& v) const { return Vector3(d[0] * v.d[0], d[1] * v.d[1], d[2] * v.d[2]); }
	const Vector3& operator*=(const Vector3& v) { d[0] *= v.d[0]; d[1] *= v.d[1]; d[2] *= v.d[2]; return *this; }
	Vector3 operator*(T s) const { return Vector3(d[0] * s, d[1] * s, d[2] * s); }
	friend Vector3 operator*(T s, const Vector3& v) { return Vector3(v.d[0] * s, v.d[1] * s, v.d[2] * s); }
	const Vector3& operator*=(T s) { d[0] *= s; d[1] *= s; d[2] *= s; return *this; }
	bool operator==(const Vector3& v) const {
		return equal<T>(d[0], v.d[0], static_cast<T>(1e-3)) && equal<T>(d[1], v.d[1], static_cast<T>(1e-3)) &&
		       equal<T>(d[2], v.d[2], static_cast<T>(1e-3));
	}
	bool operator!=(const Vector3& v) const { return !(*this == v); }
	T operator[](unsigned i) const { return d[i > 2 ? 2 : i]; }
private:
	T d[3];
};
typedef Vector3<double> Vector3d;
}  // namespace math
}  // namespace ignition
