// Stand-in for ignition/math/Line3.hh: only the type name is needed by headers on the compiled path.
#pragma once
#include <ignition/math/Vector3.hh>
namespace ignition {
namespace math {
template <typename T>
class Line3 {
public:
	Line3() {}
	Line3(const Vector3<T>& a, const Vector3<T>& b) : pts{a, b} {}
	Vector3<T> operator[](unsigned i) const { return pts[i > 1 ? 1 : i]; }
private:
	Vector3<T> pts[2];
};
typedef Line3<double> Line3d;
}  // namespace math
}  // namespace ignition
