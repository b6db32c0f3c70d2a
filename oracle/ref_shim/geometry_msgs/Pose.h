// Stand-in for the geometry_msgs message structs (plain data carriers, no arithmetic).
#pragma once
#include <string>
namespace std_msgs_shim {
struct Header {
	unsigned int seq = 0;
	double stamp = 0.0;
	std::string frame_id;
};
}  // namespace std_msgs_shim
namespace geometry_msgs {
struct Point {
	double x = 0, y = 0, z = 0;
};
struct Vector3 {
	double x = 0, y = 0, z = 0;
};
struct Quaternion {
	double x = 0, y = 0, z = 0, w = 0;
};
struct Pose {
	Point position;
	Quaternion orientation;
};
struct PoseStamped {
	std_msgs_shim::Header header;
	Pose pose;
};
struct Transform {
	Vector3 translation;
	Quaternion rotation;
};
struct TransformStamped {
	std_msgs_shim::Header header;
	std::string child_frame_id;
	Transform transform;
};
struct Twist {
	Vector3 linear;
	Vector3 angular;
};
}  // namespace geometry_msgs
