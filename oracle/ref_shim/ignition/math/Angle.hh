// Stand-in for ignition/math/Angle.hh (ignition-math 4). Normalize() = atan2(sin, cos).
#pragma once
#include <ignition/math/Helpers.hh>
namespace ignition {
namespace math {
class Angle {
public:
	Angle() : value(0.0) {}
	Angle(double radian) : value(radian) {}
	void Radian(double radian) { value = radian; }
	void Degree(double degree) { value = degree * IGN_PI / 180.0; }
	double Radian() const { return value; }
	double Degree() const { return value * 180.0 / IGN_PI; }
	void Normalize() { value = std::atan2(std::sin(value), std::cos(value)); }
	double operator*() const { return value; }
	Angle operator-(const Angle& a) const { return Angle(value - a.value); }
	Angle operator+(const Angle& a) const { return Angle(value + a.value); }
	Angle operator*(const Angle& a) const { return Angle(value * a.value); }
	Angle operator/(const Angle& a) const { return Angle(value / a.value); }
	Angle operator-=(const Angle& a) { value -= a.value; return *this; }
	Angle operator+=(const Angle& a) { value += a.value; return *this; }
	Angle operator*=(const Angle& a) { value *= a.value; return *this; }
	Angle operator/=(const Angle& a) { value /= a.value; return *this; }
	bool operator==(const Angle& a) const { return equal(value, a.value, 0.001); }
	bool operator!=(const Angle& a) const { return !(*this == a); }
	bool operator<(const Angle& a) const { return value < a.value; }
	bool operator<=(const Angle& a) const { return value < a.value || equal(value, a.value); }
	bool operator>(const Angle& a) const { return value > a.value; }
	bool operator>=(const Angle& a) const { return value > a.value || equal(value, a.value); }
private:
	double value;
};
}  // namespace math
}  // namespace ignition
