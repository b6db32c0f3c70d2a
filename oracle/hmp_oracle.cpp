/*
 * hmp_oracle.cpp -- CPU restatement (FP64, single thread per call) of the trajectory sampling +
 * scoring hot path of rayvburn/humap_local_planner.
 *
 * THIS IS TEST INFRASTRUCTURE. It is the checker the CUDA path is compared against; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call it. The
 * product (humap_local_planner_b200/) never links, loads or falls back to this file.
 *
 * Every function cites the reference file:line it restates (paths relative to the reference root).
 * Third-party arithmetic whose source is NOT vendored in the reference is marked [RECALLED]:
 *   - ignition-math 4 (Angle::Normalize, Vector3::Normalize, Quaternion Euler<->yaw)   -- pinned
 *     indirectly by the reference's test/test_geometry_*.cpp, reproduced in tests/.
 *   - base_local_planner / costmap_2d (ROS navigation, melodic/noetic): Costmap2D::worldToMap,
 *     CostmapModel::footprintCost + LineIterator, MapGrid wave front, SimpleScoredSamplingPlanner,
 *     PreferForwardCostFunction                                                          -- parity unpinned
 *   - social_nav_utils (rayvburn, no version pinned in package.xml:27,42): Gaussians and the four
 *     social cost formulations                                                           -- parity unpinned
 *   - fuzzylite 6 (HEAD at clone time, README.md:49-57): Mamdani engine                  -- parity unpinned
 * "parity unpinned" = no reference test pins numbers at that boundary; the formulation below is the
 * published / recalled one and the CUDA path is held to THIS formulation.
 */
#include "hmp_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <queue>
#include <string>
#include <array>
#include <vector>

#ifndef HMP_ORACLE_COUNT
typedef double orc_f64;   // the plain IEEE double, also in the counting build (hmp_oracle_count.cpp), where `double` is a macro
#endif

namespace {

constexpr double PI = 3.14159265358979323846;  // IGN_PI
inline double dtor(double deg) { return deg * PI / 180.0; }  // IGN_DTOR

// ------------------------------------------------------------------------------------------------
// L1 geometry: include/humap_local_planner/geometry/{angle,vector,pose}.h, src/geometry/*.cpp
// ------------------------------------------------------------------------------------------------
struct V3 {
	double x = 0, y = 0, z = 0;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(double s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }

// ignition::math::Angle::Normalize [RECALLED]: atan2(sin a, cos a); src/geometry/angle.cpp:6-10
inline double wrap(double a) { return std::atan2(std::sin(a), std::cos(a)); }

// Vector::calculateLength == ignition Vector3d::Length: 3-D, includes z (vector.h:50-52)
inline double len3(V3 v) { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }

// ignition Vector3::Normalize [RECALLED]: divide by length unless equal(length, 0, 1e-6)
inline V3 normalized(V3 v) {
	double d = len3(v);
	if (!(std::abs(d - 0.0) <= 1e-6)) {
		return {v.x / d, v.y / d, v.z / d};
	}
	return v;
}

// Vector::calculateDirection, src/geometry/vector.cpp:42-45 (result Angle is normalised by its ctor)
inline double direction(V3 v) {
	V3 n = normalized(v);
	return wrap(std::atan2(n.y, n.x));
}

// Angle(const Vector&): atan2(y, x) WITHOUT normalisation, src/geometry/angle.cpp:16-18
inline double angle_of(V3 v) { return std::atan2(v.y, v.x); }

// Pose stores a quaternion; yaw is read back through Euler (pose.h:72-74). For planar poses this is
// ignition Quaterniond(0, 0, yaw) -> Normalize -> Yaw() [RECALLED]: atan2(2wz, w^2 - z^2).
inline double yaw_roundtrip(double yaw) {
	double w = std::cos(yaw / 2.0);
	double z = std::sin(yaw / 2.0);
	double s = std::sqrt(w * w + z * z);
	if (std::abs(s) <= 1e-6) {
		w = 1.0;
		z = 0.0;
	} else {
		w /= s;
		z /= s;
	}
	// Euler(): normalises a copy again
	double s2 = std::sqrt(w * w + z * z);
	double cw = w / s2, cz = z / s2;
	return std::atan2(2.0 * (cw * cz), cw * cw - cz * cz);
}

struct Pose {
	double x = 0, y = 0, yaw = 0;
	Pose() = default;
	Pose(double x_, double y_, double yaw_) : x(x_), y(y_), yaw(yaw_roundtrip(yaw_)) {}
};

// src/utils/transformations.cpp:32-45
inline Pose computeNextPose(const Pose& pose, V3 vel, double dt) {
	double new_x = pose.x + vel.x * dt;
	double new_y = pose.y + vel.y * dt;
	double new_yaw = pose.yaw + vel.z * dt;
	return Pose(new_x, new_y, wrap(new_yaw));
}

// src/utils/transformations.cpp:16-30
inline Pose computeNextPoseBaseVel(const Pose& pose, V3 vel, double dt) {
	double new_x = pose.x + (vel.x * std::cos(pose.yaw) + vel.y * std::cos(M_PI_2 + pose.yaw)) * dt;
	double new_y = pose.y + (vel.x * std::sin(pose.yaw) + vel.y * std::sin(M_PI_2 + pose.yaw)) * dt;
	double new_yaw = pose.yaw + vel.z * dt;
	return Pose(new_x, new_y, wrap(new_yaw));
}

// src/utils/transformations.cpp:319-339
inline Pose subtractPoses(const Pose& a, const Pose& b) { return Pose(a.x - b.x, a.y - b.y, a.yaw - b.yaw); }
inline Pose addPoses(const Pose& a, const Pose& b) { return Pose(a.x + b.x, a.y + b.y, a.yaw + b.yaw); }

// src/utils/transformations.cpp:128-144
inline V3 computeVelocityGlobal(V3 vl, const Pose& pose) {
	double yaw = pose.yaw;
	return {vl.x * std::cos(yaw) - vl.y * std::sin(yaw), vl.x * std::sin(yaw) + vl.y * std::cos(yaw), vl.z};
}

// src/utils/transformations.cpp:146-175
inline V3 computeVelocityLocal(V3 vg, const Pose& pose, bool holonomic = false) {
	double yaw = pose.yaw;
	if (holonomic) {
		return {vg.x * std::cos(yaw) + vg.y * std::sin(yaw), -vg.x * std::sin(yaw) + vg.y * std::cos(yaw), vg.z};
	}
	return {vg.x * std::cos(yaw) + vg.y * std::sin(yaw), 0.0, vg.z};
}

// src/utils/transformations.cpp:177-186
inline V3 computeVelocityFromPoses(const Pose& p1, const Pose& p2, double dt) {
	Pose d = subtractPoses(p2, p1);
	return {d.x / dt, d.y / dt, d.yaw / dt};
}

// src/utils/transformations.cpp:188-197
inline V3 computeBaseVelocityFromPoses(const Pose& p1, const Pose& p2, double dt) {
	Pose d = subtractPoses(p2, p1);
	double th = p1.yaw;
	return {d.x / dt * std::cos(th) + d.y / dt * std::sin(th), -d.x / dt * std::sin(th) + d.y / dt * std::cos(th),
	        d.yaw / dt};
}

// src/utils/transformations.cpp:341-389
V3 saturateVelocity(V3 cmd, double max_vel_x, double max_vel_y, double max_vel_trans, double max_vel_theta,
                    double max_vel_x_backwards) {
	double ratio_x = 1.0, ratio_omega = 1.0, ratio_y = 1.0;
	double vx = cmd.x, vy = cmd.y, omega = cmd.z;
	if (vx > max_vel_x) ratio_x = max_vel_x / vx;
	if (vy > max_vel_y || vy < -max_vel_y) ratio_y = std::abs(vy / max_vel_y);
	if (omega > max_vel_theta || omega < -max_vel_theta) ratio_omega = std::abs(max_vel_theta / omega);
	if (vx < -std::abs(max_vel_x_backwards)) ratio_x = -std::abs(max_vel_x_backwards) / vx;
	vx *= ratio_x;
	vy *= ratio_y;
	omega *= ratio_omega;
	double vel_linear = std::hypot(vx, vy);
	if (vel_linear > max_vel_trans) {
		double r = max_vel_trans / vel_linear;
		vx *= r;
		vy *= r;
	}
	return {vx, vy, omega};
}

// src/utils/transformations.cpp:61-126
V3 computeTwist(const Pose& pose, V3 force, double robot_mass, double min_vel_x, double max_vel_x, double max_rot_vel,
                double twist_rotation_compensation) {
	if (len3(force) <= 1e-08 || robot_mass <= 1e-06) {
		return {0, 0, 0};
	}
	V3 acc_v = force / robot_mass;
	V3 vel_v_new = acc_v;
	double yaw = pose.yaw;
	double vv = +std::cos(yaw) * vel_v_new.x + std::sin(yaw) * vel_v_new.y;
	double vw = -std::sin(yaw) * vel_v_new.x + std::cos(yaw) * vel_v_new.y;
	double force_dir = angle_of(force);
	double ang_z_force_diff = wrap(force_dir - yaw);
	vw += twist_rotation_compensation * ang_z_force_diff;
	return saturateVelocity({vv, 0.0, vw}, max_vel_x, 0.0, max_vel_x, max_rot_vel,
	                        (min_vel_x < 0.0) ? std::abs(min_vel_x) : 0.0);
}

// src/utils/transformations.cpp:391-450
V3 adjustTwistProportional(V3 vel, V3 cmd, double min_x, double min_y, double min_th, double max_x, double max_y,
                           double max_th) {
	auto feas = [](double vel_curr, double vel_cmd, double lim_min, double lim_max, double& diff) -> double {
		diff = vel_cmd - vel_curr;
		if (vel_cmd >= vel_curr) {
			return (lim_max - vel_curr) / diff;
		}
		return (lim_min - vel_curr) / diff;
	};
	double dx, dy, dth;
	double fx = feas(vel.x, cmd.x, min_x, max_x, dx);
	double fy = feas(vel.y, cmd.y, min_y, max_y, dy);
	double fth = feas(vel.z, cmd.z, min_th, max_th, dth);
	// std::min_element over {fx, fy, fth}: first element not greater (uses operator<; NaN never smaller)
	double fmin = fx;
	if (fy < fmin) fmin = fy;
	if (fth < fmin) fmin = fth;
	if (std::isnan(fmin) || fmin >= 1.0) {
		return {vel.x + dx, vel.y + dy, vel.z + dth};
	}
	return {vel.x + dx * fmin, vel.y + dy * fmin, vel.z + dth * fmin};
}

// src/utils/transformations.cpp:199-255
V3 adjustTwistWithAccLimits(V3 vel, double acc_x, double acc_y, double acc_th, double min_x, double min_y,
                            double min_th, double max_x, double max_y, double max_th, double dt, V3 cmd,
                            bool maintain_rate) {
	double min_x_acc = std::max(min_x, vel.x - acc_x * dt);
	double min_y_acc = std::max(min_y, vel.y - acc_y * dt);
	double min_th_acc = std::max(min_th, vel.z - acc_th * dt);
	double max_x_acc = std::min(max_x, vel.x + acc_x * dt);
	double max_y_acc = std::min(max_y, vel.y + acc_y * dt);
	double max_th_acc = std::min(max_th, vel.z + acc_th * dt);
	if (!maintain_rate) {
		return {std::min(std::max(min_x_acc, cmd.x), max_x_acc), std::min(std::max(min_y_acc, cmd.y), max_y_acc),
		        std::min(std::max(min_th_acc, cmd.z), max_th_acc)};
	}
	return adjustTwistProportional(vel, cmd, min_x_acc, min_y_acc, min_th_acc, max_x_acc, max_y_acc, max_th_acc);
}

// src/utils/transformations.cpp:257-317
V3 adjustTwistWithAccAndGoalLimits(V3 vel, double acc_x, double acc_y, double acc_th, double min_x, double min_y,
                                   double min_th, double max_x, double max_y, double max_th, double dt, V3 cmd,
                                   bool maintain_rate, double dist_to_goal) {
	double acc_lim_decel = std::hypot(acc_x, acc_y);
	double speed_init_max = std::sqrt(2.0 * acc_lim_decel * dist_to_goal);
	double vel_vector_angle = 0.0;
	if (std::abs(vel.x) >= 1e-04 || std::abs(vel.y) >= 1e-04) {
		vel_vector_angle = std::atan2(cmd.y, cmd.x);
	}
	double vx_safe = std::cos(vel_vector_angle) * speed_init_max;
	double vy_safe = std::sin(vel_vector_angle) * speed_init_max;
	double max_x_corr = std::max(std::min(max_x, vx_safe), min_x);
	double max_y_corr = std::max(std::min(max_y, vy_safe), min_y);
	return adjustTwistWithAccLimits(vel, acc_x, acc_y, acc_th, min_x, min_y, min_th, max_x_corr, max_y_corr, max_th, dt,
	                                cmd, maintain_rate);
}

// ------------------------------------------------------------------------------------------------
// social_nav_utils/gaussians.h [RECALLED, parity unpinned]
// ------------------------------------------------------------------------------------------------
// 1-D Gaussian PDF; with normalize the maximum is 1.
inline double gaussian1d(double x, double mean, double variance, bool normalize) {
	double scale = 1.0;
	if (!normalize) {
		scale = 1.0 / (std::sqrt(variance) * std::sqrt(2.0 * PI));
	}
	return scale * std::exp(-((x - mean) * (x - mean)) / (2.0 * variance));
}
// Gaussian over a circular domain: max of the lobes centred at mean, mean + 2pi, mean - 2pi.
inline double calculateGaussianAngle(double x, double mean, double variance, bool normalize = false) {
	double g1 = gaussian1d(x, mean, variance, normalize);
	double g2 = gaussian1d(x, mean + 2.0 * PI, variance, normalize);
	double g3 = gaussian1d(x, mean - 2.0 * PI, variance, normalize);
	return std::max(std::max(g1, g2), g3);
}

// src/sfm/social_force_model.cpp:207-224 (static computeFactorFOV)
double computeFactorFOV(double angle_relative, double fov, bool gaussian) {
	if (!gaussian) {
		double fov_half = fov / 2.0;
		if (angle_relative < -fov_half) {
			return (PI + angle_relative) / (PI - fov_half);
		} else if (angle_relative > fov_half) {
			return (PI - angle_relative) / (PI - fov_half);
		}
		return 1.0;
	}
	double fov_variance = std::pow(fov / 2.0, 2.0);
	return calculateGaussianAngle(angle_relative, 0.0, fov_variance);
}

// ------------------------------------------------------------------------------------------------
// L2 World: include/humap_local_planner/world.h, src/world.cpp
// ------------------------------------------------------------------------------------------------
enum RelLoc { LOC_FRONT = 0, LOC_RIGHT, LOC_LEFT, LOC_BEHIND, LOC_UNSPECIFIED };  // defines.h

struct Object {  // StaticObject + DynamicObject (world.h:24-59)
	Pose robot, object;
	V3 dist_v;
	double dist = 0;
	V3 vel;
	double speed = 0;
	double dir_beta = 0;
	int rel_loc = LOC_UNSPECIFIED;
	double rel_loc_angle = 0;
	double dist_angle = 0;
};

struct Target {
	Pose robot, object;
	V3 dist_v;
	double dist = 0;
};

struct Robot {
	Pose centroid;
	V3 vel;
	double speed = 0;
	Target target, goal;
	double heading_dir = 0;
};

const double RELATIVE_LOCATION_FRONT_THRESHOLD = 9.0 * PI / 180.0;  // world.h:97
constexpr double SPEED_THRESHOLD_STATIONARY_ROBOT = 0.01;               // world.h:104
constexpr double SPEED_THRESHOLD_STATIONARY_OBJECT = 0.035;             // world.h:110

struct World {
	Robot robot;
	std::vector<Object> obstacle_static;
	std::vector<Object> obstacle_dynamic;

	World() = default;

	// src/world.cpp:13-34
	World(const Pose& centroid, const Pose& robot_pose, V3 robot_vel, const Pose& target_pose, const Pose& goal_pose) {
		robot.centroid = centroid;
		robot.vel = robot_vel;
		V3 vxy{robot_vel.x, robot_vel.y, 0.0};
		robot.speed = len3(vxy);
		if (robot.speed <= SPEED_THRESHOLD_STATIONARY_ROBOT) {
			robot.heading_dir = wrap(centroid.yaw);
		} else {
			robot.heading_dir = direction(vxy);
		}
		robot.target = createTarget(robot_pose, target_pose);
		robot.goal = createTarget(robot_pose, goal_pose);
	}

	// src/world.cpp:148-157
	static Target createTarget(const Pose& robot_pose, const Pose& target_pose) {
		Target t;
		t.robot = robot_pose;
		t.object = target_pose;
		t.dist_v = {target_pose.x - robot_pose.x, target_pose.y - robot_pose.y, 0.0};
		t.dist = len3(t.dist_v);
		return t;
	}

	// src/world.cpp:43-63
	void addObstacle(const Pose& robot_pose_closest, const Pose& obstacle_pose_closest, V3 obstacle_vel,
	                 bool force_dynamic_type = false) {
		if (force_dynamic_type || len3(obstacle_vel) > SPEED_THRESHOLD_STATIONARY_OBJECT) {
			obstacle_dynamic.push_back(createObstacleDynamic(robot_pose_closest, obstacle_pose_closest, obstacle_vel));
		} else {
			obstacle_static.push_back(createObstacleStatic(robot_pose_closest, obstacle_pose_closest));
		}
	}

	// src/world.cpp:133-146
	static Object createObstacleStatic(const Pose& robot_pose_closest, const Pose& obstacle_pose_closest) {
		Object o;
		o.robot = robot_pose_closest;
		o.object = obstacle_pose_closest;
		o.dist_v = {o.object.x - o.robot.x, o.object.y - o.robot.y, 0.0};
		o.dist = len3(o.dist_v);
		return o;
	}

	// src/world.cpp:159-190
	static Object createObstacleDynamic(const Pose& robot_pose_closest, const Pose& obstacle_pose_closest, V3 vel) {
		Object o = createObstacleStatic(robot_pose_closest, obstacle_pose_closest);
		double robot_yaw = wrap(robot_pose_closest.yaw);
		int rel_loc;
		double angle_rel, d_angle;
		computeObjectRelativeLocation(robot_yaw, o.dist_v, rel_loc, angle_rel, d_angle);
		o.vel = vel;
		o.speed = len3({vel.x, vel.y, 0.0});
		o.dir_beta = wrap(direction(vel));
		o.rel_loc = rel_loc;
		o.rel_loc_angle = wrap(angle_rel);
		o.dist_angle = wrap(d_angle);
		return o;
	}

	// src/world.cpp:192-229. NOTE: `angle_relative` is a DIFFERENCE of two Angles, which the Angle
	// operator- does not normalise (angle.h:86-88), so the thresholds see a value in (-2pi, 2pi).
	static void computeObjectRelativeLocation(double robot_yaw, V3 d_alpha_beta, int& rel_loc, double& angle_rel,
	                                          double& d_angle) {
		double angle_d = angle_of(d_alpha_beta);
		double angle_relative = angle_d - robot_yaw;
		int loc = LOC_UNSPECIFIED;
		if (std::fabs(angle_relative) <= RELATIVE_LOCATION_FRONT_THRESHOLD ||
		    std::fabs(angle_relative) >= (PI - RELATIVE_LOCATION_FRONT_THRESHOLD)) {
			loc = LOC_FRONT;
		} else if (angle_relative <= 0.0) {
			loc = LOC_RIGHT;
		} else if (angle_relative > 0.0) {
			loc = LOC_LEFT;
		}
		rel_loc = loc;
		angle_rel = angle_relative;
		d_angle = angle_d;
	}

	// src/world.cpp:86-114
	void predict(V3 robot_vel, double sim_period) {
		Pose centroid_new = computeNextPose(robot.centroid, robot_vel, sim_period);
		Pose centroid_diff = subtractPoses(centroid_new, robot.centroid);
		Pose robot_pose_new = addPoses(robot.target.robot, centroid_diff);
		World pred(centroid_new, robot_pose_new, robot_vel, robot.target.object, robot.goal.object);
		for (auto& ob : obstacle_dynamic) {
			Pose robot_new = addPoses(ob.robot, centroid_diff);
			Pose obstacle_new = computeNextPose(ob.object, ob.vel, sim_period);
			pred.addObstacle(robot_new, obstacle_new, ob.vel);
		}
		for (auto& ob : obstacle_static) {
			Pose robot_new = addPoses(ob.robot, centroid_diff);
			pred.addObstacle(robot_new, ob.object, V3{0, 0, 0});
		}
		*this = pred;
	}

	// world.h:243-256
	static double distanceClosest(const std::vector<Object>& objs) {
		if (objs.empty()) return std::numeric_limits<double>::max();
		double m = objs[0].dist;
		for (auto& o : objs) m = std::min(m, o.dist);  // min_element by operator< on dist
		return m;
	}
};

// ------------------------------------------------------------------------------------------------
// L2 Social force model: src/sfm/social_force_model.cpp
// ------------------------------------------------------------------------------------------------
struct SfmState {
	// float members, sfm/social_force_model.h:422-446
	float relaxation_time = 0, speed_desired = 0;
	float An = 0, Bn = 0, Cn = 0, Ap = 0, Bp = 0, Cp = 0, Aw = 0, Bw = 0;
	V3 force_internal, force_static, force_dynamic, force_combined;
};

// src/sfm/social_force_model.cpp:311-334
V3 computeInternalForce(V3 vel_robot, V3 d_robot_object, double mass, double speed_desired, double relaxation_time) {
	V3 to_goal_direction = normalized(d_robot_object);
	V3 ideal_vel_vector = speed_desired * to_goal_direction;
	V3 f_alpha = (mass * (1 / relaxation_time)) * (ideal_vel_vector - vel_robot);
	f_alpha.z = 0.0;
	return f_alpha;
}

// src/sfm/social_force_model.cpp:533-551
double computeThetaAlphaBetaAngle2014(V3 n_alpha, V3 d_alpha_beta) {
	return wrap(angle_of(n_alpha) - angle_of(d_alpha_beta));
}

// src/sfm/social_force_model.cpp:518-529 (not on the default path: param_description_ is 2014)
double computeThetaAlphaBetaAngle2011(V3 robot_vel, V3 object_vel) {
	double dot = robot_vel.x * object_vel.x + robot_vel.y * object_vel.y + robot_vel.z * object_vel.z;
	double cos_angle = dot / (len3(robot_vel) * len3(object_vel));
	return wrap(std::acos(cos_angle));
}

// src/sfm/social_force_model.cpp:583-613; description: 0 = 2014 (default), 1 = 2011
V3 computeNormalAlphaDirection(double robot_yaw, int description) {
	double yaw_norm = wrap(robot_yaw);
	if (description == 1) {
		yaw_norm = wrap(yaw_norm - wrap(PI));
	}
	double a = wrap(yaw_norm);  // Vector(Angle) rotates (1,0,0) by angle.normalized()
	return {std::cos(a), std::sin(a), 0.0};
}

inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// src/sfm/social_force_model.cpp:621-711
V3 computePerpendicularToNormal(V3 n_alpha, int rel_loc, int description) {
	const double sqr_zero = 1e-06 * 1e-06;
	V3 to_cross{0, 0, 0};
	if (rel_loc == LOC_LEFT) {
		to_cross = (description == 1) ? V3{0, 0, -1.0} : V3{0, 0, 1.0};
	} else if (rel_loc == LOC_RIGHT) {
		to_cross = (description == 1) ? V3{0, 0, 1.0} : V3{0, 0, -1.0};
	}
	V3 p_alpha = cross(n_alpha, to_cross);
	if (len3(p_alpha) < sqr_zero) {
		to_cross = {0, -1, 0};
		p_alpha = cross(p_alpha, to_cross);
	}
	return p_alpha;
}

// src/sfm/social_force_model.cpp:735-741
double computeRelativeSpeed(V3 actor_vel, V3 object_vel) {
	actor_vel.z = 0.0;
	object_vel.z = 0.0;
	return len3(object_vel - actor_vel);
}

// src/sfm/social_force_model.cpp:338-436 (dynamic object)
V3 interactionForceDynamic(const SfmState& s, const HmpSfm& cfg, const Robot& robot, const Object& object) {
	V3 n_alpha = computeNormalAlphaDirection(robot.centroid.yaw, 0);
	V3 f{0, 0, 0};
	if (object.dist > 7.5) return f;
	double fov_factor = computeFactorFOV(object.rel_loc_angle, 2.0 * cfg.fov, cfg.fov_factor_method == 0);
	double v_rel = computeRelativeSpeed(robot.vel, object.vel);
	if (std::abs(v_rel) < 1e-06) return f;
	double th = computeThetaAlphaBetaAngle2014(n_alpha, object.dist_v);
	V3 p_alpha = computePerpendicularToNormal(n_alpha, object.rel_loc, 0);
	double exp_normal = ((-s.Bn * th * th) / v_rel) - s.Cn * object.dist;
	double exp_perpendicular = ((-s.Bp * std::abs(th)) / v_rel) - s.Cp * object.dist;
	V3 n_scaled = n_alpha * s.An * std::exp(exp_normal);
	V3 p_scaled = p_alpha * s.Ap * std::exp(exp_perpendicular);
	n_scaled = n_scaled * fov_factor;
	p_scaled = p_scaled * fov_factor;
	return n_scaled + p_scaled;
}

// src/sfm/social_force_model.cpp:440-514 (static object, elliptical formulation)
V3 interactionForceStatic(const SfmState& s, const HmpSfm& cfg, const Robot& robot, const Object& object, double dt) {
	V3 d_alpha_i = -object.dist_v;
	d_alpha_i.z = 0.0;
	double d_len = object.dist;
	V3 y_alpha_i = robot.vel * dt;
	y_alpha_i.z = 0.0;
	V3 dy = d_alpha_i - y_alpha_i;
	double dy_len = len3(dy);
	double w = 0.5 * std::sqrt(std::pow((d_len + dy_len), 2) - std::pow(len3(y_alpha_i), 2));
	bool w_small = std::abs(w) < 1e-08;
	bool w_nan = std::isnan(w);
	bool d_short = d_len < 1e-08;
	if (w_small || w_nan || d_short) return V3{0, 0, 0};
	// NOTE the precedence: ((d + |d - y|) / 2 * w), SURVEY App. A #2
	V3 f = (s.Aw * std::exp(-w / s.Bw) * ((d_len + dy_len) / 2 * w) * 0.5) * (normalized(d_alpha_i) + normalized(dy));
	double angle_relative = wrap(direction(object.dist_v) - robot.heading_dir);
	double fov_factor = computeFactorFOV(angle_relative, 2.0 * cfg.fov, cfg.fov_factor_method == 0);
	return f * fov_factor;
}

// src/sfm/social_force_model.cpp:756-881 (only the live V4 branch)
void applyNonlinearOperations(SfmState& s, const HmpSfm& cfg) {
	double mag = len3(s.force_combined);
	if (mag >= cfg.max_force) {
		double c = cfg.max_force / mag;
		s.force_internal = s.force_internal * c;
		s.force_dynamic = s.force_dynamic * c;
		s.force_static = s.force_static * c;
		s.force_combined = s.force_internal + s.force_dynamic + s.force_static;
	} else if (mag <= cfg.min_force) {
		double extension_len = std::fabs(mag - cfg.min_force);
		V3 extension = extension_len * normalized(s.force_combined);
		s.force_dynamic = s.force_dynamic + extension;
		s.force_combined = s.force_internal + s.force_dynamic + s.force_static;
		// sf_values_{1, 0} shift register only alters force_combined_, never read by the generator (App. A #15)
	}
}

// src/sfm/social_force_model.cpp:134-202
void computeSocialForce(SfmState& s, const HmpSfm& cfg, const World& world, double dt) {
	s.force_internal = s.force_static = s.force_dynamic = s.force_combined = V3{0, 0, 0};
	const Robot& robot = world.robot;
	s.force_internal = computeInternalForce(robot.vel, robot.target.dist_v, cfg.mass, s.speed_desired, s.relaxation_time);
	if (!cfg.disable_interaction_forces) {
		for (const Object& o : world.obstacle_static) {
			s.force_static = s.force_static + interactionForceStatic(s, cfg, robot, o, dt);
		}
		for (const Object& o : world.obstacle_dynamic) {
			s.force_dynamic = s.force_dynamic + interactionForceDynamic(s, cfg, robot, o);
		}
	}
	// factorInForceCoefficients :745-752
	s.force_internal = s.force_internal * cfg.internal_force_factor;
	s.force_static = s.force_static * cfg.static_interaction_force_factor;
	s.force_dynamic = s.force_dynamic * cfg.dynamic_interaction_force_factor;
	s.force_combined = s.force_internal + s.force_dynamic + s.force_static;
	if (cfg.filter_forces) {
		applyNonlinearOperations(s, cfg);
	}
}

// ------------------------------------------------------------------------------------------------
// L2 Fuzzy inference: src/fuzz/{processor,trapezoid_parted,trapezoid_loc_dep,trapezoid_loc_indep,
// social_conductor}.cpp over fuzzylite 6 [RECALLED, parity unpinned]
// ------------------------------------------------------------------------------------------------
constexpr double FL_MACHEPS = 1e-6;  // fl::fuzzylite::macheps()
inline bool fl_isEq(double a, double b) { return a == b || std::abs(a - b) < FL_MACHEPS || (a != a && b != b); }
inline bool fl_isLt(double a, double b) { return !fl_isEq(a, b) && a < b; }
inline bool fl_isLE(double a, double b) { return fl_isEq(a, b) || a < b; }
inline bool fl_isGt(double a, double b) { return !fl_isEq(a, b) && a > b; }

struct FlTerm {  // fl::Trapezoid (4 vertices) or fl::Triangle (c == d unused, is_triangle)
	double a, b, c, d;
	double height;
	bool triangle;
};

// fl::Trapezoid::membership / fl::Triangle::membership [RECALLED]
double fl_membership(const FlTerm& t, double x) {
	if (std::isnan(x)) return std::numeric_limits<double>::quiet_NaN();
	if (t.triangle) {
		if (fl_isLt(x, t.a) || fl_isGt(x, t.c)) return t.height * 0.0;
		if (fl_isEq(x, t.b)) return t.height * 1.0;
		if (fl_isLt(x, t.b)) {
			if (t.a == -std::numeric_limits<double>::infinity()) return t.height * 1.0;
			return t.height * (x - t.a) / (t.b - t.a);
		}
		if (t.c == std::numeric_limits<double>::infinity()) return t.height * 1.0;
		return t.height * (t.c - x) / (t.c - t.b);
	}
	if (fl_isLt(x, t.a) || fl_isGt(x, t.d)) return t.height * 0.0;
	if (fl_isLt(x, t.b)) return t.height * std::min(1.0, (x - t.a) / (t.b - t.a));
	if (fl_isLE(x, t.c)) return t.height * 1.0;
	if (fl_isLt(x, t.d)) return t.height * (t.d - x) / (t.d - t.c);
	if (t.d == std::numeric_limits<double>::infinity()) return t.height * 1.0;
	return t.height * 0.0;
}

inline FlTerm trapezoid_deg(double a, double b, double c, double d) { return {dtor(a), dtor(b), dtor(c), dtor(d), 1.0, false}; }
inline FlTerm triangle_deg(double a, double b, double c) { return {dtor(a), dtor(b), dtor(c), 0.0, 1.0, true}; }

// std::to_string(double) -> "%f" (6 decimals) -> fuzzylite Op::toScalar; src/fuzz/trapezoid_parted.cpp:199-212
inline double quantize6(double v) {
	if (std::isnan(v)) return v;
	char buf[64];
	std::snprintf(buf, sizeof(buf), "%f", (orc_f64)v);
	return std::strtod(buf, nullptr);
}

struct TrapezoidParted {  // two fl::Trapezoid terms "<name>A", "<name>B"
	FlTerm t[2];
};

const double FL_NAN = std::numeric_limits<double>::quiet_NaN();

// src/fuzz/trapezoid_parted.cpp:57-189
bool trapezoidPartedUpdate(TrapezoidParted& tp, double intersection, double start, double end) {
	const double RANGE_EXTENSION = dtor(5);
	double p[4] = {FL_NAN, FL_NAN, FL_NAN, FL_NAN};
	double pw[4] = {FL_NAN, FL_NAN, FL_NAN, FL_NAN};
	bool have_p = false, have_pw = false;
	bool status = false;
	double a = wrap(start - intersection);
	double d = wrap(end + intersection);
	auto gen = [](double* out, double v0, double v1, double v2, double v3) {
		out[0] = quantize6(v0);
		out[1] = quantize6(v1);
		out[2] = quantize6(v2);
		out[3] = quantize6(v3);
	};
	if (a < d) {
		if (a >= (-PI) && d <= (+PI)) {
			gen(p, a, start, end, d);
			have_p = true;
		}
		// else: reference throws std::runtime_error (unreachable: a, d are normalised)
	} else {
		if (a > start) {
			// CASE 1
			double wrap_angle_n = wrap(-PI - start);
			double b_out_range = PI - wrap_angle_n;
			gen(p, a, b_out_range, b_out_range + RANGE_EXTENSION, b_out_range + RANGE_EXTENSION);
			double wrap_angle_w = wrap(PI - a);
			double a_out_range = -PI - wrap_angle_w;
			double len_raw = 2 * intersection + (end - start);
			if (len_raw > (2 * PI)) {
				d = end + intersection;  // Angle(end + intersection_, false)
			}
			gen(pw, a_out_range, start, end, d);
			have_p = have_pw = true;
		} else if (start >= end) {
			// CASE 2
			gen(p, a, start, +PI, +PI);
			gen(pw, -PI, -PI, end, d);
			have_p = have_pw = true;
		} else if (end >= d) {
			// CASE 3
			double wr = std::fabs(-PI - d);
			double d_out_range = PI + wr;
			gen(p, a, start, end, d_out_range);
			wr = std::fabs(PI - end);
			double c_out_range = -PI - wr;
			gen(pw, c_out_range - RANGE_EXTENSION, c_out_range - RANGE_EXTENSION, c_out_range, d);
			have_p = have_pw = true;
		}
		status = true;
	}
	(void)have_p;
	(void)have_pw;
	// configure(params): 5 values incl. height 1.0 when generated; "nan nan nan nan" otherwise
	tp.t[0] = {p[0], p[1], p[2], p[3], 1.0, false};
	tp.t[1] = {pw[0], pw[1], pw[2], pw[3], status ? 1.0 : 0.0, false};
	return status;
}

struct FisOutput {
	double value = 0, membership = 0;
	int term = -1;  // index into output terms, -1 "none"
};

enum FisDir { DIR_OUTWARDS = 0, DIR_CROSS_FRONT, DIR_CROSS_BEHIND, DIR_EQUAL, DIR_OPPOSITE };
enum FisLoc { LOCT_BACK_A = 0, LOCT_BACK_RIGHT, LOCT_FRONT_RIGHT, LOCT_FRONT, LOCT_FRONT_LEFT, LOCT_BACK_LEFT, LOCT_BACK_B };
enum FisOut {
	OUT_ACCELERATE = 0, OUT_TURN_RIGHT_ACCELERATE, OUT_TURN_RIGHT, OUT_TURN_RIGHT_DECELERATE, OUT_DECELERATE_A,
	OUT_STOP_A, OUT_DECELERATE_B, OUT_STOP_B, OUT_TURN_LEFT_DECELERATE, OUT_TURN_LEFT, OUT_TURN_LEFT_ACCELERATE,
	OUT_COUNT
};

struct FisRule {
	int loc, dir, out;
};

// src/fuzz/processor.cpp:148-171
const FisRule FIS_RULES[18] = {
    {LOCT_FRONT, DIR_OPPOSITE, OUT_TURN_RIGHT},
    {LOCT_FRONT, DIR_OUTWARDS, OUT_DECELERATE_A},
    {LOCT_FRONT, DIR_OUTWARDS, OUT_DECELERATE_B},
    {LOCT_FRONT, DIR_EQUAL, OUT_DECELERATE_A},
    {LOCT_FRONT, DIR_EQUAL, OUT_DECELERATE_B},
    {LOCT_FRONT, DIR_CROSS_FRONT, OUT_TURN_RIGHT},
    {LOCT_FRONT_RIGHT, DIR_CROSS_BEHIND, OUT_TURN_LEFT},
    {LOCT_FRONT_RIGHT, DIR_OPPOSITE, OUT_TURN_LEFT},
    {LOCT_FRONT_RIGHT, DIR_OUTWARDS, OUT_TURN_LEFT},
    {LOCT_FRONT_RIGHT, DIR_EQUAL, OUT_TURN_LEFT_ACCELERATE},
    {LOCT_FRONT_RIGHT, DIR_CROSS_FRONT, OUT_TURN_RIGHT},
    {LOCT_BACK_RIGHT, DIR_CROSS_BEHIND, OUT_TURN_LEFT_ACCELERATE},
    {LOCT_BACK_RIGHT, DIR_OPPOSITE, OUT_TURN_RIGHT},
    {LOCT_BACK_RIGHT, DIR_EQUAL, OUT_TURN_RIGHT_ACCELERATE},
    {LOCT_BACK_RIGHT, DIR_CROSS_FRONT, OUT_ACCELERATE},
    {LOCT_BACK_LEFT, DIR_CROSS_BEHIND, OUT_ACCELERATE},
    {LOCT_FRONT_LEFT, DIR_CROSS_BEHIND, OUT_TURN_RIGHT},
    {LOCT_FRONT_LEFT, DIR_CROSS_FRONT, OUT_TURN_RIGHT_ACCELERATE},
};

struct FisEngine {
	FlTerm loc_terms[7];
	FlTerm out_terms[OUT_COUNT];
	TrapezoidParted dir_terms[5];

	// src/fuzz/processor.cpp:19-114
	FisEngine() {
		loc_terms[LOCT_BACK_A] = triangle_deg(-180, -180, -160);
		loc_terms[LOCT_BACK_RIGHT] = trapezoid_deg(-180, -150, -120, -90);
		loc_terms[LOCT_FRONT_RIGHT] = trapezoid_deg(-120, -90, -30, 0);
		loc_terms[LOCT_FRONT] = triangle_deg(-20, 0, 20);
		loc_terms[LOCT_FRONT_LEFT] = trapezoid_deg(0, 30, 90, 120);
		loc_terms[LOCT_BACK_LEFT] = trapezoid_deg(90, 120, 150, 180);
		loc_terms[LOCT_BACK_B] = triangle_deg(160, 180, 180);
		out_terms[OUT_ACCELERATE] = trapezoid_deg(-30, -15, -15, 30);
		out_terms[OUT_TURN_RIGHT_ACCELERATE] = trapezoid_deg(-75, -60, -30, -15);
		out_terms[OUT_TURN_RIGHT] = trapezoid_deg(-120, -105, -75, -60);
		out_terms[OUT_TURN_RIGHT_DECELERATE] = trapezoid_deg(-155, -140, -120, -105);
		out_terms[OUT_DECELERATE_A] = trapezoid_deg(-180, -165, -155, -140);
		out_terms[OUT_STOP_A] = triangle_deg(-195, -180, -165);
		out_terms[OUT_DECELERATE_B] = trapezoid_deg(140, 155, 165, 180);
		out_terms[OUT_STOP_B] = triangle_deg(165, 180, 195);
		out_terms[OUT_TURN_LEFT_DECELERATE] = trapezoid_deg(105, 120, 140, 155);
		out_terms[OUT_TURN_LEFT] = trapezoid_deg(60, 75, 105, 120);
		out_terms[OUT_TURN_LEFT_ACCELERATE] = trapezoid_deg(15, 30, 60, 75);
		for (auto& d : dir_terms) {
			d.t[0] = {FL_NAN, FL_NAN, FL_NAN, FL_NAN, 1.0, false};
			d.t[1] = {FL_NAN, FL_NAN, FL_NAN, FL_NAN, 1.0, false};
		}
	}

	// src/fuzz/processor.cpp:335-394 + trapezoid_loc_dep.cpp:19-35 + trapezoid_loc_indep.cpp:15-27
	void updateRegions(double alpha_dir, double d_alpha_beta_angle, double rel_loc) {
		const double I = dtor(10);
		double gamma_eq = wrap(alpha_dir);
		double gamma_opp = wrap(gamma_eq + PI);
		double gamma_cc = wrap(d_alpha_beta_angle + PI);
		bool right = rel_loc < 0.0;  // decodeRelativeLocation :383-394 (>= 0 -> LEFT)
		auto loc_dep = [&](TrapezoidParted& tp, double gs, double ge) {
			if (right) {
				trapezoidPartedUpdate(tp, I, gs, ge);
			} else {
				trapezoidPartedUpdate(tp, I, ge, gs);
			}
		};
		loc_dep(dir_terms[DIR_OUTWARDS], gamma_opp, gamma_eq);
		loc_dep(dir_terms[DIR_CROSS_FRONT], gamma_eq, gamma_cc);
		loc_dep(dir_terms[DIR_CROSS_BEHIND], gamma_cc, gamma_opp);
		const double interval = dtor(std::max(20.0, 1e-03)) / 2.0;
		trapezoidPartedUpdate(dir_terms[DIR_EQUAL], I, wrap(gamma_eq - interval), wrap(gamma_eq + interval));
		trapezoidPartedUpdate(dir_terms[DIR_OPPOSITE], I, wrap(gamma_opp - interval), wrap(gamma_opp + interval));
	}

	// one iteration of Processor::process, src/fuzz/processor.cpp:214-267
	FisOutput processOne(double dir_alpha, double dir_beta, double rel_loc, double dist_angle) {
		auto bound = [](double v) { return v > PI ? PI : (v < -PI ? -PI : v); };  // setLockValueInRange
		double location = bound(rel_loc);
		updateRegions(dir_alpha, dist_angle, rel_loc);
		double dirv = bound(wrap(dir_beta));
		// antecedents
		double mu_loc[7], mu_dir[5];
		for (int i = 0; i < 7; ++i) mu_loc[i] = fl_membership(loc_terms[i], location);
		for (int i = 0; i < 5; ++i) {
			double a = fl_membership(dir_terms[i].t[0], dirv);
			double b = fl_membership(dir_terms[i].t[1], dirv);
			mu_dir[i] = a + b - (a * b);  // fl::AlgebraicSum
		}
		// General activation: Minimum conjunction, AlgebraicProduct implication; triggered if degree > 0 (macheps)
		double deg[18];
		bool trig[18];
		bool any = false;
		for (int r = 0; r < 18; ++r) {
			deg[r] = 1.0 * std::min(mu_loc[FIS_RULES[r].loc], mu_dir[FIS_RULES[r].dir]);
			trig[r] = fl_isGt(deg[r], 0.0);
			any = any || trig[r];
		}
		FisOutput out;
		double value = FL_NAN;  // default value fl::nan
		if (any) {
			// fl::Centroid(100) over [-pi, pi], Aggregated with fl::Maximum
			const int resolution = 100;
			const double dx = (PI - (-PI)) / resolution;
			double area = 0, xcentroid = 0;
			for (int i = 0; i < resolution; ++i) {
				double x = -PI + (i + 0.5) * dx;
				double mu = 0.0;
				for (int r = 0; r < 18; ++r) {
					if (!trig[r]) continue;
					double m = fl_membership(out_terms[FIS_RULES[r].out], x) * deg[r];
					mu = std::max(mu, m);
				}
				xcentroid += mu * x;
				area += mu;
			}
			value = xcentroid / area;
		}
		if (!std::isnan(value)) value = bound(value);
		// highestMembership
		int best = -1;
		double ymax = 0.0;
		for (int k = 0; k < OUT_COUNT; ++k) {
			double y = fl_membership(out_terms[k], value);
			if (fl_isGt(y, ymax)) {
				ymax = y;
				best = k;
			}
		}
		if (best < 0) {
			out.value = 0.0;
			out.membership = 0.0;
			out.term = -1;
			return out;
		}
		out.value = value;
		out.membership = ymax;
		out.term = best;
		return out;
	}
};

// src/fuzz/social_conductor.cpp:162-179
double computeBehaviourStrengthExponential(double action_range, double dist_to_agent, double speed_agent,
                                           double speed_obstacle) {
	if (dist_to_agent > action_range) return 0.0;
	double mutual_speed = speed_agent + speed_obstacle;
	double speed_factor = std::exp(+mutual_speed) - 1.0;
	double dist_factor = std::exp(-dist_to_agent);
	return speed_factor * dist_factor;
}

// src/fuzz/social_conductor.cpp:37-105, :181-190
V3 computeBehaviourForce(const HmpFis& cfg, double As, const Pose& pose_agent, double speed_agent,
                         const std::vector<FisOutput>& fis, const std::vector<double>& speeds,
                         const std::vector<double>& dists, const std::vector<double>& rel_locs) {
	V3 force{0, 0, 0};
	for (size_t i = 0; i < dists.size(); ++i) {
		if (fis[i].membership <= 0.0) continue;
		double a = wrap(wrap(fis[i].value));
		V3 v_temp{std::cos(a), std::sin(a), 0.0};
		double strength = computeBehaviourStrengthExponential(cfg.human_action_range, dists[i], speed_agent, speeds[i]);
		double fov_factor = 1.0;
		if (cfg.fov_factor_method == 0) {
			fov_factor = computeFactorFOV(rel_locs[i], cfg.fov, true);
		} else if (cfg.fov_factor_method == 1) {
			fov_factor = computeFactorFOV(rel_locs[i], cfg.fov, false);
		}
		double magnitude = As * fis[i].membership * strength * fov_factor;
		force = force + (v_temp * magnitude);
	}
	// Vector::rotate(double), src/geometry/vector.cpp:56-64
	double yaw = pose_agent.yaw;
	V3 r{force.x * std::cos(yaw) - force.y * std::sin(yaw), force.x * std::sin(yaw) + force.y * std::cos(yaw), force.z};
	return r * cfg.force_factor;
}

// ------------------------------------------------------------------------------------------------
// Trajectory wrappers: base_local_planner::Trajectory [RECALLED] and include/humap_local_planner/trajectory.h
// ------------------------------------------------------------------------------------------------
struct BlpTrajectory {
	double xv = 0, yv = 0, thetav = 0, cost = 0, time_delta = 0;
	std::vector<double> x, y, th;
	size_t size() const { return x.size(); }
};

struct Traj {  // humap_local_planner::Trajectory
	double dt = 0;
	std::vector<Pose> poses;
	std::vector<V3> vels;
};

// trajectory.h:43-103
Traj makeTrajectory(const BlpTrajectory& t, bool convert_to_global_velocities) {
	Traj out;
	out.dt = t.time_delta;
	for (size_t i = 0; i < t.size(); ++i) {
		Pose pose_curr(t.x[i], t.y[i], t.th[i]);
		if (i == 0) {
			out.poses.push_back(pose_curr);
			V3 velocity{t.xv, t.yv, t.thetav};
			if (convert_to_global_velocities) velocity = computeVelocityGlobal(velocity, pose_curr);
			out.vels.push_back(velocity);
			continue;
		}
		if (i == 1) {
			out.poses.push_back(pose_curr);
			continue;
		}
		Pose pose_prev(t.x[i - 1], t.y[i - 1], t.th[i - 1]);
		V3 vel = convert_to_global_velocities ? computeVelocityFromPoses(pose_prev, pose_curr, out.dt)
		                                      : computeBaseVelocityFromPoses(pose_prev, pose_curr, out.dt);
		out.poses.push_back(pose_curr);
		out.vels.push_back(vel);
	}
	return out;
}

// trajectory.h:160-193 (people: constant velocity) and :198-222 (groups: static, dt 1e-3)
Traj predictObject(const Pose& pose0, V3 vel_global, double dt, unsigned steps) {
	Traj out;
	out.dt = dt;
	out.poses.push_back(pose0);
	out.vels.push_back(vel_global);
	for (unsigned i = 1; i < steps; ++i) {
		out.poses.push_back(computeNextPose(out.poses.back(), vel_global, dt));
		if (i == steps - 1) continue;
		out.vels.push_back(vel_global);
	}
	return out;
}

// ------------------------------------------------------------------------------------------------
// costmap_2d / base_local_planner restatements [RECALLED, parity unpinned]
// ------------------------------------------------------------------------------------------------
struct Costmap {
	const uint8_t* cells = nullptr;
	int size_x = 0, size_y = 0;
	double origin_x = 0, origin_y = 0, resolution = 1;
	// Costmap2D::worldToMap
	bool worldToMap(double wx, double wy, unsigned& mx, unsigned& my) const {
		if (wx < origin_x || wy < origin_y) return false;
		mx = (int)((wx - origin_x) / resolution);
		my = (int)((wy - origin_y) / resolution);
		if (mx < (unsigned)size_x && my < (unsigned)size_y) return true;
		return false;
	}
	void mapToWorld(unsigned mx, unsigned my, double& wx, double& wy) const {
		wx = origin_x + (mx + 0.5) * resolution;
		wy = origin_y + (my + 0.5) * resolution;
	}
	uint8_t getCost(unsigned mx, unsigned my) const { return cells[(size_t)my * size_x + mx]; }
};
constexpr uint8_t NO_INFORMATION = 255, LETHAL_OBSTACLE = 254, INSCRIBED_INFLATED_OBSTACLE = 253;

// CostmapModel::pointCost / lineCost (LineIterator = Bresenham) / footprintCost
double pointCost(const Costmap& cm, int x, int y) {
	uint8_t cost = cm.getCost(x, y);
	if (cost == NO_INFORMATION) return -2;
	if (cost == LETHAL_OBSTACLE) return -1;
	return cost;
}

double lineCost(const Costmap& cm, int x0, int x1, int y0, int y1) {
	double line_cost = 0.0;
	int deltax = std::abs(x1 - x0), deltay = std::abs(y1 - y0);
	int x = x0, y = y0;
	int xinc1, xinc2, yinc1, yinc2, den, num, numadd, numpixels;
	if (x1 >= x0) { xinc1 = 1; xinc2 = 1; } else { xinc1 = -1; xinc2 = -1; }
	if (y1 >= y0) { yinc1 = 1; yinc2 = 1; } else { yinc1 = -1; yinc2 = -1; }
	if (deltax >= deltay) {
		xinc1 = 0; yinc2 = 0; den = deltax; num = deltax / 2; numadd = deltay; numpixels = deltax;
	} else {
		xinc2 = 0; yinc1 = 0; den = deltay; num = deltay / 2; numadd = deltax; numpixels = deltay;
	}
	for (int curpixel = 0; curpixel <= numpixels; ++curpixel) {
		double pc = pointCost(cm, x, y);
		if (pc < 0) return pc;
		if (line_cost < pc) line_cost = pc;
		num += numadd;
		if (num >= den) {
			num -= den;
			x += xinc1;
			y += yinc1;
		}
		x += xinc2;
		y += yinc2;
	}
	return line_cost;
}

// WorldModel::footprintCost(x, y, theta, spec) -> CostmapModel::footprintCost(position, oriented footprint)
double worldModelFootprintCost(const Costmap& cm, double x, double y, double theta, const std::vector<double>& spec) {
	const size_t n = spec.size() / 2;
	double cos_th = std::cos(theta), sin_th = std::sin(theta);
	std::vector<double> fx(n), fy(n);
	for (size_t i = 0; i < n; ++i) {
		fx[i] = x + (spec[2 * i] * cos_th - spec[2 * i + 1] * sin_th);
		fy[i] = y + (spec[2 * i] * sin_th + spec[2 * i + 1] * cos_th);
	}
	unsigned cell_x, cell_y;
	if (!cm.worldToMap(x, y, cell_x, cell_y)) return -3.0;
	if (n < 3) {
		uint8_t cost = cm.getCost(cell_x, cell_y);
		if (cost == NO_INFORMATION) return -2.0;
		if (cost == LETHAL_OBSTACLE || cost == INSCRIBED_INFLATED_OBSTACLE) return -1.0;
		return cost;
	}
	unsigned x0, x1, y0, y1;
	double line_cost = 0.0, footprint_cost = 0.0;
	for (size_t i = 0; i < n - 1; ++i) {
		if (!cm.worldToMap(fx[i], fy[i], x0, y0)) return -3.0;
		if (!cm.worldToMap(fx[i + 1], fy[i + 1], x1, y1)) return -3.0;
		line_cost = lineCost(cm, x0, x1, y0, y1);
		footprint_cost = std::max(line_cost, footprint_cost);
		if (line_cost < 0) return line_cost;
	}
	if (!cm.worldToMap(fx[n - 1], fy[n - 1], x0, y0)) return -3.0;
	if (!cm.worldToMap(fx[0], fy[0], x1, y1)) return -3.0;
	line_cost = lineCost(cm, x0, x1, y0, y1);
	footprint_cost = std::max(line_cost, footprint_cost);
	if (line_cost < 0) return line_cost;
	return footprint_cost;
}

// src/obstacle_separation_cost_function.cpp:164-242
double obstacleFootprintCost(const Costmap& cm, double x, double y, double th, double separation_dist, int kernel,
                             const std::vector<double>& spec) {
	double footprint_cost = worldModelFootprintCost(cm, x, y, th, spec);
	bool without_sep = std::abs(separation_dist) < 1e-03;
	if (!without_sep) {
		auto fk = [&](double offset, double angle) {
			double xk = x + offset * std::cos(th + angle);
			double yk = y + offset * std::sin(th + angle);
			return worldModelFootprintCost(cm, xk, yk, th, spec);
		};
		std::vector<double> costs;
		costs.push_back(footprint_cost);
		if (kernel == 0) {
			costs.push_back(fk(separation_dist, 0.0));
			costs.push_back(fk(separation_dist, +M_PI_2));
			costs.push_back(fk(separation_dist, +M_PI));
			costs.push_back(fk(separation_dist, -M_PI_2));
		} else if (kernel == 1) {
			costs.push_back(fk(separation_dist, 0.0));
			costs.push_back(fk(separation_dist, +M_PI_4));
			costs.push_back(fk(separation_dist, +M_PI_2));
			costs.push_back(fk(separation_dist, +3.0 * M_PI_4));
			costs.push_back(fk(separation_dist, +M_PI));
			costs.push_back(fk(separation_dist, -3.0 * M_PI_4));
			costs.push_back(fk(separation_dist, -M_PI_2));
			costs.push_back(fk(separation_dist, -M_PI_4));
		}
		double max_cost = *std::max_element(costs.begin(), costs.end());
		double min_cost = *std::min_element(costs.begin(), costs.end());
		footprint_cost = (min_cost < 0.0) ? min_cost : max_cost;
	}
	if (footprint_cost < 0) return -6.0;
	unsigned cell_x, cell_y;
	if (!cm.worldToMap(x, y, cell_x, cell_y)) return -7.0;
	return std::max(std::max(0.0, footprint_cost), double(cm.getCost(cell_x, cell_y)));
}

// src/obstacle_separation_cost_function.cpp:85-114
double scoreObstacle(const Costmap& cm, const BlpTrajectory& traj, const HmpCosts& c, const std::vector<double>& spec) {
	double cost = 0;
	if (spec.empty()) return -9;
	for (size_t i = 0; i < traj.size(); ++i) {
		double f_cost = obstacleFootprintCost(cm, traj.x[i], traj.y[i], traj.th[i], c.occdist_separation,
		                                      c.occdist_separation_kernel, spec);
		if (f_cost < 0) return f_cost;
		if (c.occdist_sum_scores) {
			cost += f_cost;
		} else {
			cost = std::max(cost, f_cost);
		}
	}
	return cost;
}

// MapGrid wave front [RECALLED]: resetPathDist + setTargetCells / setLocalGoal + computeTargetDistance
void mapgridCompute(const Costmap& cm, const double* plan_xy, int n_plan, bool local_goal, double* target_dist) {
	const int sx = cm.size_x, sy = cm.size_y;
	const double obstacle_costs = (double)sx * sy;
	const double unreachable = (double)sx * sy + 1;
	const size_t N = (size_t)sx * sy;
	std::vector<char> mark(N, 0);
	for (size_t i = 0; i < N; ++i) target_dist[i] = unreachable;
	if (n_plan <= 0) return;
	// adjustPlanResolution
	std::vector<double> px, py;
	double last_x = plan_xy[0], last_y = plan_xy[1];
	px.push_back(last_x);
	py.push_back(last_y);
	double min_sq_resolution = cm.resolution * cm.resolution;
	for (int i = 1; i < n_plan; ++i) {
		double loop_x = plan_xy[2 * i], loop_y = plan_xy[2 * i + 1];
		double sqdist = (loop_x - last_x) * (loop_x - last_x) + (loop_y - last_y) * (loop_y - last_y);
		if (sqdist > min_sq_resolution) {
			int steps = (int)std::ceil((std::sqrt(sqdist)) / cm.resolution);
			double deltax = (loop_x - last_x) / steps;
			double deltay = (loop_y - last_y) / steps;
			for (int j = 1; j < steps; ++j) {
				px.push_back(last_x + j * deltax);
				py.push_back(last_y + j * deltay);
			}
		}
		px.push_back(loop_x);
		py.push_back(loop_y);
		last_x = loop_x;
		last_y = loop_y;
	}
	std::queue<size_t> q;
	bool started_path = false;
	if (!local_goal) {
		for (size_t i = 0; i < px.size(); ++i) {
			unsigned mx, my;
			if (cm.worldToMap(px[i], py[i], mx, my) && cm.getCost(mx, my) != NO_INFORMATION) {
				size_t idx = (size_t)my * sx + mx;
				target_dist[idx] = 0.0;
				mark[idx] = 1;
				q.push(idx);
				started_path = true;
			} else if (started_path) {
				break;
			}
		}
		if (!started_path) return;
	} else {
		int gx = -1, gy = -1;
		for (size_t i = 0; i < px.size(); ++i) {
			unsigned mx, my;
			if (cm.worldToMap(px[i], py[i], mx, my) && cm.getCost(mx, my) != NO_INFORMATION) {
				gx = mx;
				gy = my;
				started_path = true;
			} else if (started_path) {
				break;
			}
		}
		if (!started_path) return;
		if (gx >= 0 && gy >= 0) {
			size_t idx = (size_t)gy * sx + gx;
			target_dist[idx] = 0.0;
			mark[idx] = 1;
			q.push(idx);
		}
	}
	auto update = [&](size_t cur, size_t chk) -> bool {
		uint8_t cost = cm.cells[chk];
		if (cost == LETHAL_OBSTACLE || cost == INSCRIBED_INFLATED_OBSTACLE || cost == NO_INFORMATION) {
			target_dist[chk] = obstacle_costs;
			return false;
		}
		double nd = target_dist[cur] + 1;
		if (nd < target_dist[chk]) target_dist[chk] = nd;
		return true;
	};
	while (!q.empty()) {
		size_t cur = q.front();
		q.pop();
		unsigned cx = cur % sx, cy = cur / sx;
		auto visit = [&](size_t chk) {
			if (!mark[chk]) {
				mark[chk] = 1;
				if (update(cur, chk)) q.push(chk);
			}
		};
		if (cx > 0) visit(cur - 1);
		if (cx < (unsigned)sx - 1) visit(cur + 1);
		if (cy > 0) visit(cur - sx);
		if (cy < (unsigned)sy - 1) visit(cur + sx);
	}
}

struct MapGridCritic {
	const double* target_dist = nullptr;
	int size_x = 0, size_y = 0;
	double xshift = 0, yshift = 0;
	bool stop_on_failure = false;
	int n_kernel_size = 0;
	double n_cost_multiplier = 3.0;
	double highest_valid_cost_prev = 0, highest_valid_cost = 0;
	double obstacleCosts() const { return (double)size_x * size_y; }
	double unreachableCellCosts() const { return (double)size_x * size_y + 1; }

	// MapGrid::operator()(unsigned x, unsigned y) = map_[size_x_ * y + x] with unsigned wrap-around: an
	// index is read iff it lands inside the array (where the reference's behaviour is defined); the
	// out-of-bounds reads of the reference are skipped.
	bool cell(long long x, long long y, double& v) const {
		long long idx = (long long)size_x * y + x;
		if (idx < 0 || idx >= (long long)size_x * size_y) return false;
		v = target_dist[idx];
		return true;
	}

	// src/map_grid_cost_function.cpp:81-140
	double getCellCosts(unsigned px, unsigned py) {
		double grid_dist = target_dist[(size_t)py * size_x + px];
		if (grid_dist != unreachableCellCosts() || n_kernel_size <= 0) {
			if (grid_dist != obstacleCosts()) {
				highest_valid_cost = std::max(highest_valid_cost, grid_dist);
			}
			return grid_dist;
		}
		std::vector<double> npts = {grid_dist};
		int offset = 1;
		while (offset <= n_kernel_size) {
			bool px_poff = ((long long)px + offset) <= size_x;
			bool py_poff = ((long long)py + offset) <= size_y;
			const bool px_noff = true, py_noff = true;  // unsigned (px - offset) >= 0
			double v;
			long long X = px, Y = py, o = offset;
			if (px_poff && py_poff && cell(X + o, Y + o, v)) npts.push_back(v);
			if (px_poff && cell(X + o, Y, v)) npts.push_back(v);
			if (py_poff && cell(X, Y + o, v)) npts.push_back(v);
			if (py_noff && cell(X, Y - o, v)) npts.push_back(v);
			if (px_noff && cell(X - o, Y, v)) npts.push_back(v);
			if (px_noff && py_noff && cell(X - o, Y - o, v)) npts.push_back(v);
			if (px_poff && py_noff && cell(X + o, Y - o, v)) npts.push_back(v);
			if (px_noff && py_poff && cell(X - o, Y + o, v)) npts.push_back(v);
			offset++;
		}
		double max_cost = *std::max_element(npts.begin(), npts.end());
		if (max_cost != obstacleCosts() && max_cost != unreachableCellCosts()) {
			double min_cost = *std::min_element(npts.begin(), npts.end()) * n_cost_multiplier;
			highest_valid_cost = std::max(highest_valid_cost, min_cost);
			return min_cost;
		}
		return highest_valid_cost_prev;
	}

	// src/map_grid_cost_function.cpp:142-196 (aggregation `Last`, the only one HumapPlanner constructs)
	double score(const Costmap& cm, const BlpTrajectory& traj) {
		double cost = 0.0;
		for (size_t i = 0; i < traj.size(); ++i) {
			double px = traj.x[i], py = traj.y[i], pth = traj.th[i];
			if (xshift != 0.0) {
				px = px + xshift * std::cos(pth);
				py = py + xshift * std::sin(pth);
			}
			if (yshift != 0.0) {
				px = px + yshift * std::cos(pth + M_PI_2);
				py = py + yshift * std::sin(pth + M_PI_2);
			}
			unsigned cell_x, cell_y;
			if (!cm.worldToMap(px, py, cell_x, cell_y)) return -4.0;
			double grid_dist = getCellCosts(cell_x, cell_y);
			if (stop_on_failure) {
				if (grid_dist == obstacleCosts()) return -3.0;
				if (grid_dist == unreachableCellCosts()) return -2.0;
			}
			cost = grid_dist;
		}
		return cost;
	}
};

// ------------------------------------------------------------------------------------------------
// social_nav_utils cost formulations [RECALLED / restated from the published descriptions; parity unpinned]
// See DESIGN.md "third-party formulations" for the statement of each.
// ------------------------------------------------------------------------------------------------
// exp(-1/2 d^T Sigma^-1 d) for Sigma = [[a, b1], [b2, c]] (Gaussian PDF divided by its maximum).
inline double gaussian2dNormalized(double dx, double dy, double a, double b1, double b2, double c) {
	double det = a * c - b1 * b2;
	double q = (c * dx * dx - (b1 + b2) * dx * dy + a * dy * dy) / det;
	return std::exp(-0.5 * q);
}

// social_nav_utils::PersonalSpaceIntrusion(...).normalize().getScale(): Kirby's asymmetric Gaussian
// (front / rear variance chosen by the side of the person the robot is on, side variance across),
// rotated to the person's heading, plus the person's position covariance; normalised to peak 1.
double personalSpaceIntrusion(double xp, double yp, double yawp, double cxx, double cxy, double cyx, double cyy,
                              double var_front, double var_rear, double var_side, double xr, double yr) {
	double dx = xr - xp, dy = yr - yp;
	double c = std::cos(yawp), s = std::sin(yawp);
	double along = dx * c + dy * s;  // robot position along the person's heading
	double var_h = (along >= 0.0) ? var_front : var_rear;
	double a = var_h * c * c + var_side * s * s + cxx;
	double b = (var_h - var_side) * c * s;
	double cc = var_h * s * s + var_side * c * c + cyy;
	return gaussian2dNormalized(dx, dy, a, b + cxy, b + cyx, cc);
}

// social_nav_utils::FormationSpaceIntrusion(...).normalize().getScale(): O-space Gaussian with the given
// variances along the group's axes, rotated by the group's yaw, plus the group's position covariance.
double formationSpaceIntrusion(double xg, double yg, double yawg, double var_x, double var_y, double cxx, double cxy,
                               double cyy, double xr, double yr) {
	double dx = xr - xg, dy = yr - yg;
	double c = std::cos(yawg), s = std::sin(yawg);
	double a = var_x * c * c + var_y * s * s + cxx;
	double b = (var_x - var_y) * c * s + cxy;
	double cc = var_x * s * s + var_y * c * c + cyy;
	return gaussian2dNormalized(dx, dy, a, b, b, cc);
}

// social_nav_utils::HeadingDirectionDisturbance(...).normalize(r_robot, v_max).getScale():
// product of (i) a circular Gaussian of the robot's motion direction around the direction that crosses
// the person's centre, its standard deviation being the half-angle under which the person's occupancy
// circle (inflated by the position uncertainty) is seen from the robot, (ii) a circular Gaussian of the
// robot's bearing in the person's frame with sigma = half of the person's FOV, (iii) speed / max_speed,
// (iv) (r_robot + r_person) / max(distance, r_robot + r_person).
double headingDirectionDisturbance(double xp, double yp, double yawp, double cxx, double cxy, double cyy, double xr,
                                   double yr, double yawr, double vxr, double vyr, double person_radius, double fov_person,
                                   double robot_circumradius, double max_speed) {
	(void)cxy;
	(void)yawr;
	double dx = xr - xp, dy = yr - yp;
	double dist = std::sqrt(dx * dx + dy * dy);
	double speed = std::sqrt(vxr * vxr + vyr * vyr);
	if (speed < 1e-9 || dist < 1e-9) return 0.0;
	double dist_angle = std::atan2(dy, dx);
	double rel_loc = wrap(dist_angle - yawp);
	double gamma_cc = wrap(dist_angle + PI);
	double motion_dir = std::atan2(vyr, vxr);
	double radius_eff = person_radius + std::sqrt(0.5 * (cxx + cyy));
	double half_angle = std::atan2(radius_eff, dist);
	double var_dir = half_angle * half_angle;
	double g_dir = calculateGaussianAngle(wrap(motion_dir - gamma_cc), 0.0, var_dir, true);
	double var_fov = (fov_person / 2.0) * (fov_person / 2.0);
	double g_fov = calculateGaussianAngle(rel_loc, 0.0, var_fov, true);
	double d_min = robot_circumradius + person_radius;
	double dist_factor = d_min / std::max(dist, d_min);
	double speed_factor = speed / max_speed;
	return g_dir * g_fov * speed_factor * dist_factor;
}

// social_nav_utils::PassingSpeedComfort(distance, speed, d_min, v_max).getDiscomfortNormalized():
// discomfort grows linearly with the robot speed (saturating at v_max) and decays exponentially with
// the clearance beyond d_min (1 m decay length).
double passingSpeedDiscomfort(double distance, double speed, double min_dist, double max_speed) {
	double sp = std::min(std::max(speed / max_speed, 0.0), 1.0);
	double clearance = std::max(distance - min_dist, 0.0);
	return sp * std::exp(-clearance);
}

// ------------------------------------------------------------------------------------------------
// First-party critics on the wrapped Trajectory
// ------------------------------------------------------------------------------------------------
struct PersonPred {
	HmpPerson p;
	Traj traj;
};
struct GroupPred {
	HmpGroup g;
	Traj traj;
};

struct PlanState {
	const HmpParams* P = nullptr;
	Costmap cm;
	MapGridCritic grids[HMP_NUM_MAPGRIDS];
	std::vector<double> footprint;
	World world;
	V3 vel_local;  // vel_
	std::vector<PersonPred> people;
	std::vector<GroupPred> groups;
};

// src/ttc_cost_function.cpp:29-137,181-188
double scoreTTC(const PlanState& st, const BlpTrajectory& traj) {
	const HmpCosts& c = st.P->costs;
	if (traj.size() == 0) return -10.0;
	const double dt = traj.time_delta;
	if (dt <= 1e-12 || std::isinf(dt) || std::isnan(dt)) return -12.0;
	Traj robot_traj = makeTrajectory(traj, true);
	// World::predict(const Trajectory&), src/world.cpp:116-131
	std::vector<World> seq;
	seq.push_back(st.world);
	for (const V3& vel : robot_traj.vels) {
		World w = seq.back();
		w.predict(vel, robot_traj.dt);
		seq.push_back(w);
	}
	auto cost = [](double ttc, double total) {
		if (ttc <= 0.0) ttc = 1e-04;
		return total / ttc;
	};
	double timestamp = 0.0;
	double traj_time = traj.size() * dt;
	double dmin_s = std::numeric_limits<double>::max(), dmin_d = std::numeric_limits<double>::max();
	for (const World& w : seq) {
		dmin_s = std::min(dmin_s, World::distanceClosest(w.obstacle_static));
		dmin_d = std::min(dmin_d, World::distanceClosest(w.obstacle_dynamic));
		if (dmin_s <= c.ttc_collision_distance || dmin_d <= c.ttc_collision_distance) {
			return cost(timestamp, traj_time + c.ttc_rollout_time);
		}
		timestamp += dt;
	}
	V3 robot_global_vel = robot_traj.vels.back();
	for (double t = 0; t < c.ttc_rollout_time; t += dt) {
		World w = seq.back();
		dmin_s = std::min(dmin_s, World::distanceClosest(w.obstacle_static));
		dmin_d = std::min(dmin_d, World::distanceClosest(w.obstacle_dynamic));
		if (dmin_s <= c.ttc_collision_distance || dmin_d <= c.ttc_collision_distance) {
			return cost(timestamp, traj_time + c.ttc_rollout_time);
		}
		w.predict(robot_global_vel, dt);
		timestamp += dt;
		seq.push_back(w);
	}
	return 0.0;
}

// src/unsaturated_translation_cost_function.cpp:31-87
double scoreUnsaturated(const PlanState& st, const BlpTrajectory& traj) {
	const HmpCosts& c = st.P->costs;
	if (traj.size() == 0) return 0.0;
	Traj t = makeTrajectory(traj, false);
	if (t.vels.empty()) return 0.0;
	double vx_dev = 0, vy_dev = 0, vxy_dev = 0;
	size_t n = 0;
	vx_dev += std::abs(t.vels[0].x - c.unsat_max_vel_x);
	vy_dev += std::abs(t.vels[0].y - c.unsat_max_vel_y);
	vxy_dev += std::abs(std::hypot(t.vels[0].x, t.vels[0].y) - c.unsat_max_trans_vel);
	n++;
	if (c.unsat_whole_horizon) {
		for (size_t i = 1; i < t.vels.size(); ++i) {
			vx_dev += std::abs(t.vels[i].x - c.unsat_max_vel_x);
			vy_dev += std::abs(t.vels[i].y - c.unsat_max_vel_y);
			vxy_dev += std::abs(std::hypot(t.vels[i].x, t.vels[i].y) - c.unsat_max_trans_vel);
			n++;
		}
	}
	return std::max(std::max(vx_dev, vy_dev), vxy_dev) / static_cast<double>(n);
}

// base_local_planner::PreferForwardCostFunction::scoreTrajectory [RECALLED]
double scoreBackward(const PlanState& st, const BlpTrajectory& traj) {
	if (traj.xv < 0.0) return st.P->costs.backward_penalty;
	if (traj.xv < 0.1 && std::fabs(traj.thetav) < 0.2) return st.P->costs.backward_penalty;
	return std::fabs(traj.thetav) * 10;
}

// src/heading_change_smoothness_cost_function.cpp:15-43
double scoreHeadingChange(const PlanState& st, const BlpTrajectory& traj) {
	if (traj.size() == 0) return 0.0;
	Traj t = makeTrajectory(traj, false);
	if (t.vels.empty()) return 0.0;
	double hcs = std::abs(t.vels[0].z - st.vel_local.z);
	for (size_t i = 1; i < t.vels.size(); ++i) {
		double domega = t.vels[i].z - t.vels[i - 1].z;
		hcs += (std::abs(domega) / t.dt);
	}
	return hcs / static_cast<double>(t.vels.size() + 1);
}

// src/velocity_smoothness_cost_function.cpp:18-50
double scoreVelocitySmoothness(const PlanState& st, const BlpTrajectory& traj) {
	if (traj.size() == 0) return 0.0;
	Traj t = makeTrajectory(traj, false);
	if (t.vels.empty()) return 0.0;
	double vx_dev = std::abs(t.vels[0].x - st.vel_local.x);
	double vy_dev = std::abs(t.vels[0].y - st.vel_local.y);
	for (size_t i = 1; i < t.vels.size(); ++i) {
		vx_dev += std::abs(t.vels[i].x - t.vels[i - 1].x);
		vy_dev += std::abs(t.vels[i].y - t.vels[i - 1].y);
	}
	return (vx_dev + vy_dev) / static_cast<double>(t.vels.size() + 1);
}

// src/heading_disturbance_cost_function.cpp:35-96
double scoreHeadingDisturbance(const PlanState& st, const BlpTrajectory& traj) {
	const HmpCosts& c = st.P->costs;
	if (st.people.empty()) return 0.0;
	Traj rt = makeTrajectory(traj, true);
	unsigned i_end_whole = (unsigned)rt.vels.size();
	unsigned i_end = c.hd_whole_horizon ? i_end_whole : std::min(i_end_whole, 1u);
	double overall = -std::numeric_limits<double>::infinity();
	for (const PersonPred& pp : st.people) {
		double best = -std::numeric_limits<double>::infinity();
		for (unsigned i = 0; i < i_end; ++i) {
			const Pose& pr = rt.poses.at(i);
			const V3& vr = rt.vels.at(i);
			const Pose& p = pp.traj.poses.at(i);
			double v = headingDirectionDisturbance(p.x, p.y, p.yaw, pp.p.cov_xx, pp.p.cov_xy, pp.p.cov_yy, pr.x, pr.y, pr.yaw,
			                                       vr.x, vr.y, c.hd_person_model_radius, c.hd_fov_person,
			                                       c.hd_robot_circumradius, c.hd_max_speed);
			best = std::max(best, v);
		}
		overall = std::max(overall, best);
	}
	return overall;
}

// src/personal_space_intrusion_cost_function.cpp:21-89
double scorePersonalSpace(const PlanState& st, const BlpTrajectory& traj) {
	const HmpCosts& c = st.P->costs;
	if (st.people.empty()) return 0.0;
	Traj rt = makeTrajectory(traj, true);
	unsigned i_end_whole = (unsigned)rt.vels.size();
	unsigned i_end = c.psi_whole_horizon ? i_end_whole : std::min(i_end_whole, 1u);
	double overall = -std::numeric_limits<double>::infinity();
	for (const PersonPred& pp : st.people) {
		double best = -std::numeric_limits<double>::infinity();
		for (unsigned i = 0; i < i_end; ++i) {
			const Pose& pr = rt.poses.at(i);
			const Pose& p = pp.traj.poses.at(i);
			const V3& vp = pp.traj.vels.at(i);
			double vel_lin = std::hypot(vp.x, vp.y);
			double var_front = std::max(2.0 * vel_lin, 0.5);
			double var_side = (2.0 / 3.0) * var_front;
			double var_rear = (1.0 / 2.0) * var_front;
			double v = personalSpaceIntrusion(p.x, p.y, p.yaw, pp.p.cov_xx, pp.p.cov_xy, pp.p.cov_yx, pp.p.cov_yy, var_front,
			                                  var_rear, var_side, pr.x, pr.y);
			best = std::max(best, v);
		}
		overall = std::max(overall, best);
	}
	return overall;
}

// src/fformation_space_intrusion_cost_function.cpp:21-88
double scoreFformation(const PlanState& st, const BlpTrajectory& traj) {
	const HmpCosts& c = st.P->costs;
	if (st.groups.empty()) return 0.0;
	Traj rt = makeTrajectory(traj, true);
	unsigned i_end_whole = (unsigned)rt.poses.size();
	unsigned i_end = c.fsi_whole_horizon ? i_end_whole : std::min(i_end_whole, 1u);
	double overall = -std::numeric_limits<double>::infinity();
	for (const GroupPred& gp : st.groups) {
		double best = -std::numeric_limits<double>::infinity();
		for (unsigned i = 0; i < i_end; ++i) {
			const Pose& pr = rt.poses.at(i);
			const Pose& pg = gp.traj.poses.at(i);
			double var_x = std::pow((gp.g.span_x / 2.0) / 2.0, 2);
			double var_y = std::pow((gp.g.span_y / 2.0) / 2.0, 2);
			double v = formationSpaceIntrusion(pg.x, pg.y, pg.yaw, var_x, var_y, gp.g.cov_xx, gp.g.cov_xy, gp.g.cov_yy, pr.x,
			                                   pr.y);
			best = std::max(best, v);
		}
		overall = std::max(overall, best);
	}
	return overall;
}

// src/passing_speed_cost_function.cpp:25-80
double scorePassingSpeed(const PlanState& st, const BlpTrajectory& traj) {
	const HmpCosts& c = st.P->costs;
	if (st.people.empty()) return 0.0;
	Traj rt = makeTrajectory(traj, true);
	unsigned i_end_whole = (unsigned)rt.vels.size();
	unsigned i_end = c.ps_whole_horizon ? i_end_whole : std::min(i_end_whole, 1u);
	double overall = -std::numeric_limits<double>::infinity();
	for (const PersonPred& pp : st.people) {
		double best = -std::numeric_limits<double>::infinity();
		for (unsigned i = 0; i < i_end; ++i) {
			const Pose& pr = rt.poses.at(i);
			const V3& vr = rt.vels.at(i);
			const Pose& p = pp.traj.poses.at(i);
			double distance = std::sqrt(std::pow(pr.x - p.x, 2) + std::pow(pr.y - p.y, 2));
			double speed_robot = std::hypot(vr.x, vr.y);
			best = std::max(best, passingSpeedDiscomfort(distance, speed_robot, c.ps_min_dist, c.ps_max_speed));
		}
		overall = std::max(overall, best);
	}
	return overall;
}

double scoreCritic(PlanState& st, int k, const BlpTrajectory& traj) {
	switch (k) {
		case HMP_COST_OBSTACLE: return scoreObstacle(st.cm, traj, st.P->costs, st.footprint);
		case HMP_COST_PATH: return st.grids[HMP_GRID_PATH].score(st.cm, traj);
		case HMP_COST_GOAL: return st.grids[HMP_GRID_GOAL].score(st.cm, traj);
		case HMP_COST_ALIGNMENT: return st.grids[HMP_GRID_ALIGNMENT].score(st.cm, traj);
		case HMP_COST_GOAL_FRONT: return st.grids[HMP_GRID_GOAL_FRONT].score(st.cm, traj);
		case HMP_COST_UNSATURATED: return scoreUnsaturated(st, traj);
		case HMP_COST_BACKWARD: return scoreBackward(st, traj);
		case HMP_COST_TTC: return scoreTTC(st, traj);
		case HMP_COST_HEADING_CHANGE: return scoreHeadingChange(st, traj);
		case HMP_COST_VEL_SMOOTHNESS: return scoreVelocitySmoothness(st, traj);
		case HMP_COST_HEADING_DIST: return scoreHeadingDisturbance(st, traj);
		case HMP_COST_PERSONAL_SPACE: return scorePersonalSpace(st, traj);
		case HMP_COST_FFORMATION: return scoreFformation(st, traj);
		case HMP_COST_PASSING_SPEED: return scorePassingSpeed(st, traj);
	}
	return 0.0;
}

// SimpleScoredSamplingPlanner::scoreTrajectory [RECALLED]; raw[] receives the raw critic outputs (NaN = not evaluated)
double scoreTrajectoryAll(PlanState& st, const BlpTrajectory& traj, double best_traj_cost, bool early_exit, double* raw) {
	double traj_cost = 0;
	for (int k = 0; k < HMP_NUM_COSTS; ++k) raw[k] = std::numeric_limits<double>::quiet_NaN();
	for (int k = 0; k < HMP_NUM_COSTS; ++k) {
		double scale = st.P->costs.scale[k];
		if (scale == 0) continue;
		double cost = scoreCritic(st, k, traj);
		raw[k] = cost;
		if (cost < 0) {
			traj_cost = cost;
			break;
		}
		if (cost != 0) cost *= scale;
		traj_cost += cost;
		if (early_exit && best_traj_cost > 0) {
			if (traj_cost > best_traj_cost) break;
		}
	}
	return traj_cost;
}

// ------------------------------------------------------------------------------------------------
// L3 SocialTrajectoryGenerator: src/social_trajectory_generator.cpp
// ------------------------------------------------------------------------------------------------
// :465-498
std::vector<double> computeAmplifierSamples(double amp_min, double amp_max, double granularity) {
	std::vector<double> samples;
	int num_amps = (int)std::ceil((amp_max - amp_min) / granularity);
	for (int i = 0; i <= num_amps; i++) {
		double v = amp_min + granularity * i;
		if (v > amp_max) {
			samples.push_back(amp_max);
			break;
		}
		samples.push_back(v);
	}
	if (samples.empty()) samples.push_back(0.0);
	return samples;
}

// :166-217 nested loops, speed outermost ... As innermost
std::vector<HmpSample> buildSamples(const HmpSampling& s, const HmpSample* extra, int n_extra) {
	std::vector<std::vector<double>> lists(HMP_NUM_AMPLIFIERS);
	size_t total = 1;
	for (int a = 0; a < HMP_NUM_AMPLIFIERS; ++a) {
		lists[a] = computeAmplifierSamples(s.amp_min[a], s.amp_max[a], s.amp_granularity[a]);
		total *= lists[a].size();
	}
	std::vector<HmpSample> out;
	out.reserve(total + n_extra);
	for (size_t idx = 0; idx < total; ++idx) {
		HmpSample smp;
		size_t rem = idx;
		for (int a = HMP_NUM_AMPLIFIERS - 1; a >= 0; --a) {
			size_t n = lists[a].size();
			smp.amp[a] = (orc_f64)lists[a][rem % n];
			rem /= n;
		}
		out.push_back(smp);
	}
	for (int i = 0; i < n_extra; ++i) out.push_back(extra[i]);
	return out;
}

// :584-599
int computeStepsNumber(const HmpGeneral& g, double speed_linear, double speed_angular) {
	if (g.discretize_by_time) {
		return std::ceil(g.sim_time / g.sim_granularity);
	}
	double sim_time_distance = speed_linear * g.sim_time;
	double sim_time_angle = std::fabs(speed_angular) * g.sim_time;
	return (int)std::ceil(std::max(sim_time_distance / g.sim_granularity, sim_time_angle / g.angular_sim_granularity));
}

// :556-582
bool areVelocityLimitsFulfilled(const HmpLimits& l, double speed_linear, double speed_angular, double eps) {
	bool vel_trans_wrong = (l.min_vel_trans >= 0) && ((speed_linear + eps) < l.min_vel_trans);
	bool vel_theta_wrong = (l.min_vel_theta >= 0) && ((std::abs(speed_angular) + eps) < l.min_vel_theta);
	if (vel_trans_wrong && vel_theta_wrong) return false;
	if ((l.max_vel_trans >= 0) && ((speed_linear - eps) > l.max_vel_trans)) return false;
	return true;
}

// ------------------------------------------------------------------------------------------------
// Environment model: HumapPlanner::createEnvironmentModel (src/humap_planner.cpp:930-1052, first-party) over the
// teb_local_planner obstacle classes [RECALLED: getClosestPoint / getMinimumDistance / toPolygonMsg of Point, Circular,
// Line and Polygon obstacles and closest_point_on_line_segment_2d; parity unpinned] and the first-party footprint models
// (include/humap_local_planner/robot_footprint_model.h:55-140).
// ------------------------------------------------------------------------------------------------
struct P2 {
	double x, y;
};
inline P2 closestPointOnSegment(P2 p, P2 s, P2 e) {
	double dx = e.x - s.x, dy = e.y - s.y;
	double sq = dx * dx + dy * dy;
	if (sq == 0) return s;
	double u = ((p.x - s.x) * dx + (p.y - s.y) * dy) / sq;
	if (u <= 0) return s;
	if (u >= 1) return e;
	return {s.x + u * dx, s.y + u * dy};
}
inline double dist2d(P2 a, P2 b) { return std::hypot(a.x - b.x, a.y - b.y); }

struct ShapeView {
	const HmpShape* s;
	const double* verts;
	P2 vertex(int i) const { return {verts[2 * (s->first_vertex + i)], verts[2 * (s->first_vertex + i) + 1]}; }
	// Obstacle::getClosestPoint(position)
	P2 closestPoint(P2 p) const {
		switch (s->type) {
			case HMP_SHAPE_POINT: return {s->x, s->y};
			case HMP_SHAPE_CIRCLE: {
				double dx = p.x - s->x, dy = p.y - s->y;
				double n = std::sqrt(dx * dx + dy * dy);
				return {s->x + s->radius * (dx / n), s->y + s->radius * (dy / n)};   // Eigen normalized(): 0/0 for a centre hit
			}
			case HMP_SHAPE_LINE: return closestPointOnSegment(p, {s->x, s->y}, {s->x2, s->y2});
			default: {
				const int n = s->n_vertices;
				if (n == 1) return vertex(0);
				P2 best = vertex(0);
				double dmin = HUGE_VAL;
				for (int i = 0; i < n - 1; ++i) {
					P2 q = closestPointOnSegment(p, vertex(i), vertex(i + 1));
					double d = dist2d(q, p);
					if (d < dmin) { dmin = d; best = q; }
				}
				if (n > 2) {
					P2 q = closestPointOnSegment(p, vertex(n - 1), vertex(0));
					double d = dist2d(q, p);
					if (d < dmin) { dmin = d; best = q; }
				}
				return best;
			}
		}
	}
	// Obstacle::getMinimumDistance(position)
	double minimumDistance(P2 p) const {
		switch (s->type) {
			case HMP_SHAPE_POINT: return dist2d(p, {s->x, s->y});
			case HMP_SHAPE_CIRCLE: return dist2d(p, {s->x, s->y}) - s->radius;
			case HMP_SHAPE_LINE: return dist2d(p, closestPointOnSegment(p, {s->x, s->y}, {s->x2, s->y2}));
			default: return dist2d(p, closestPoint(p));
		}
	}
	// points of Obstacle::toPolygonMsg
	int numPolygonPoints() const { return s->type == HMP_SHAPE_POLYGON ? s->n_vertices : (s->type == HMP_SHAPE_LINE ? 2 : 1); }
	P2 polygonPoint(int i) const {
		if (s->type == HMP_SHAPE_POLYGON) return vertex(i);
		if (s->type == HMP_SHAPE_LINE && i == 1) return {s->x2, s->y2};
		return {s->x, s->y};
	}
};

// ---- include/humap_local_planner/utils/vector_calculations.h (first-party) over teb_local_planner's
// closest_point_on_line_segment_2d / check_line_segments_intersection_2d [RECALLED, parity unpinned] -------------------------
inline double norm2(P2 v) { return std::sqrt(v.x * v.x + v.y * v.y); }
inline P2 minus(P2 a, P2 b) { return {a.x - b.x, a.y - b.y}; }
// v - v.normalized() * r (Eigen >= 3.3: normalized() returns a zero vector unchanged)
inline P2 subtractRadius(P2 v, double r) {
	double n2 = v.x * v.x + v.y * v.y;
	if (!(n2 > 0.0)) return {v.x - v.x * r, v.y - v.y * r};
	double n = std::sqrt(n2);
	return {v.x - (v.x / n) * r, v.y - (v.y / n) * r};
}
inline P2 vectorPointToSegment(P2 point, P2 s, P2 e) { return minus(point, closestPointOnSegment(point, s, e)); }
inline bool checkLineSegmentsIntersection(P2 a0, P2 a1, P2 b0, P2 b1) {
	P2 line1 = minus(a1, a0), line2 = minus(b1, b0);
	double denom = line1.x * line2.y - line2.x * line1.y;
	if (denom == 0) return false;
	bool denom_positive = denom > 0;
	P2 aux = minus(a0, b0);
	double s_numer = line1.x * aux.y - line1.y * aux.x;
	if ((s_numer < 0) == denom_positive) return false;
	double t_numer = line2.x * aux.y - line2.y * aux.x;
	if ((t_numer < 0) == denom_positive) return false;
	if (((s_numer > denom) == denom_positive) || ((t_numer > denom) == denom_positive)) return false;
	return true;
}
// vector_calculations.h:33-63. Intersecting segments: the reference returns a default-constructed Eigen::Vector2d, i.e.
// uninitialised memory; restated as the zero vector (deviation documented in DESIGN.md)
inline P2 vectorSegmentToSegment(P2 l1s, P2 l1e, P2 l2s, P2 l2e) {
	if (checkLineSegmentsIntersection(l1s, l1e, l2s, l2e)) return {0.0, 0.0};
	P2 v[4] = {vectorPointToSegment(l1s, l2s, l2e), vectorPointToSegment(l1e, l2s, l2e), vectorPointToSegment(l2s, l1s, l1e),
	           vectorPointToSegment(l2e, l1s, l1e)};
	int idx = 0;
	double shortest = norm2(v[0]);
	for (int i = 1; i < 4; ++i) {
		double len = norm2(v[i]);
		if (len < shortest) {
			shortest = len;
			idx = i;
		}
	}
	return v[idx];
}
inline P2 vectorPointToPolygon(P2 point, const std::vector<P2>& vertices) {   // :65-95
	double dist = HUGE_VAL;
	P2 vec{0.0, 0.0};
	if (vertices.size() == 1) return minus(point, vertices.front());
	for (int i = 0; i < (int)vertices.size() - 1; ++i) {
		P2 nv = vectorPointToSegment(point, vertices[i], vertices[i + 1]);
		double d = norm2(nv);
		if (d < dist) {
			dist = d;
			vec = nv;
		}
	}
	if (vertices.size() > 2) {
		P2 nv = vectorPointToSegment(point, vertices.back(), vertices.front());
		if (norm2(nv) < dist) return nv;
	}
	return vec;
}
inline P2 vectorSegmentToPolygon(P2 ls, P2 le, const std::vector<P2>& vertices) {   // :97-131
	double dist = HUGE_VAL;
	P2 vec{0.0, 0.0};
	if (vertices.size() == 1) return vectorPointToSegment(vertices.front(), ls, le);
	for (int i = 0; i < (int)vertices.size() - 1; ++i) {
		P2 nv = vectorSegmentToSegment(ls, le, vertices[i], vertices[i + 1]);
		double d = norm2(nv);
		if (d < dist) {
			dist = d;
			vec = nv;
		}
	}
	if (vertices.size() > 2) {
		P2 nv = vectorSegmentToSegment(ls, le, vertices.back(), vertices.front());
		double d = norm2(nv);
		if (d < dist) {
			dist = d;
			vec = nv;
		}
	}
	return vec;
}
inline P2 vectorPolygonToPolygon(const std::vector<P2>& v1, const std::vector<P2>& v2) {   // :133-165
	double dist = HUGE_VAL;
	P2 vec{0.0, 0.0};
	if (v1.size() == 1) return vectorPointToPolygon(v1.front(), v2);
	for (int i = 0; i < (int)v1.size() - 1; ++i) {
		P2 nv = vectorSegmentToPolygon(v1[i], v1[i + 1], v2);
		double d = norm2(nv);
		if (d < dist) {
			dist = d;
			vec = nv;
		}
	}
	if (v1.size() > 2) {
		P2 nv = vectorSegmentToPolygon(v1.back(), v1.front(), v2);
		double d = norm2(nv);
		if (d < dist) {
			dist = d;
			vec = nv;
		}
	}
	return vec;
}
inline std::vector<P2> shapeVertices(const ShapeView& sv) {
	std::vector<P2> v;
	for (int i = 0; i < sv.s->n_vertices; ++i) v.push_back(sv.vertex(i));
	return v;
}
// Obstacle::getShortestVector(line_start, line_end), include/humap_local_planner/obstacles.h:98,167,235,298
inline P2 shortestVectorToSegment(const ShapeView& sv, P2 ls, P2 le) {
	const HmpShape* s = sv.s;
	switch (s->type) {
		case HMP_SHAPE_POINT: return vectorPointToSegment({s->x, s->y}, ls, le);
		case HMP_SHAPE_CIRCLE: return subtractRadius(vectorPointToSegment({s->x, s->y}, ls, le), s->radius);
		case HMP_SHAPE_LINE: return vectorSegmentToSegment({s->x, s->y}, {s->x2, s->y2}, ls, le);
		default: return vectorSegmentToPolygon(ls, le, shapeVertices(sv));
	}
}
// Obstacle::getShortestVector(polygon), obstacles.h:101,171,238,301
inline P2 shortestVectorToPolygon(const ShapeView& sv, const std::vector<P2>& polygon) {
	const HmpShape* s = sv.s;
	switch (s->type) {
		case HMP_SHAPE_POINT: return vectorPointToPolygon({s->x, s->y}, polygon);
		case HMP_SHAPE_CIRCLE: return subtractRadius(vectorPointToPolygon({s->x, s->y}, polygon), s->radius);
		case HMP_SHAPE_LINE: return vectorSegmentToPolygon({s->x, s->y}, {s->x2, s->y2}, polygon);
		default: return vectorPolygonToPolygon(polygon, shapeVertices(sv));
	}
}

// BaseRobotFootprintModel::calculateClosestPoints of the five footprint models (robot_footprint_model.h:89-95, :132-139,
// :203-223, :279-291, :340-344), restated as written
inline void calculateClosestPoints(const HmpEnvParams& env, const double pose[3], const ShapeView& sv, Pose& robot_out, Pose& obstacle_out) {
	const P2 position{pose[0], pose[1]};
	switch (env.robot_model) {
		case HMP_ROBOT_TWO_CIRCLES: {
			const double front_offset = env.two_circles[0], front_radius = env.two_circles[1], rear_offset = env.two_circles[2],
			             rear_radius = env.two_circles[3];
			const P2 dir{std::cos(pose[2]), std::sin(pose[2])};   // PoseSE2::orientationUnitVec
			const P2 centre_front{position.x + front_offset * dir.x, position.y + front_offset * dir.y};
			const P2 centre_rear{position.x + rear_offset * dir.x, position.y + rear_offset * dir.y};
			const P2 obs_pt_front = sv.closestPoint(centre_front), obs_pt_rear = sv.closestPoint(centre_rear);
			// "robot_pt_*" are vectors (v - v.normalized() * radius), and the shortest one becomes pts.robot
			const P2 hyp[4] = {subtractRadius(minus(obs_pt_front, centre_front), front_radius), subtractRadius(minus(obs_pt_rear, centre_front), front_radius),
			                   subtractRadius(minus(obs_pt_front, centre_rear), rear_radius), subtractRadius(minus(obs_pt_rear, centre_rear), rear_radius)};
			const P2 obs[4] = {obs_pt_front, obs_pt_rear, obs_pt_front, obs_pt_rear};
			int best = 0;
			for (int k = 1; k < 4; ++k)
				if (norm2(hyp[k]) < norm2(hyp[best])) best = k;   // std::sort by norm: insertion sort for 4 elements, first minimum first
			obstacle_out = Pose(obs[best].x, obs[best].y, 0.0);
			robot_out = Pose(hyp[best].x, hyp[best].y, pose[2]);
			return;
		}
		case HMP_ROBOT_LINE: {
			const double c = std::cos(pose[2]), s = std::sin(pose[2]);   // teb LineRobotFootprint::transformToWorld
			const P2 ls{position.x + c * env.line_xy[0] - s * env.line_xy[1], position.y + s * env.line_xy[0] + c * env.line_xy[1]};
			const P2 le{position.x + c * env.line_xy[2] - s * env.line_xy[3], position.y + s * env.line_xy[2] + c * env.line_xy[3]};
			const P2 v = shortestVectorToSegment(sv, ls, le);
			robot_out = Pose(pose[0], pose[1], pose[2]);
			obstacle_out = Pose(position.x - v.x, position.y - v.y, 0.0);
			return;
		}
		case HMP_ROBOT_POLYGON: {
			std::vector<P2> vertices;   // vertices_ of the footprint: ROBOT frame, not transformed by the pose (as written, :340-344)
			for (int i = 0; i < env.n_polygon; ++i) vertices.push_back({env.polygon_xy[2 * i], env.polygon_xy[2 * i + 1]});
			const P2 v = shortestVectorToPolygon(sv, vertices);
			robot_out = Pose(pose[0], pose[1], pose[2]);
			obstacle_out = Pose(v.x, v.y, 0.0);
			return;
		}
		default: break;
	}
	const P2 obstacle_pt = sv.closestPoint(position);
	obstacle_out = Pose(obstacle_pt.x, obstacle_pt.y, 0.0);
	if (env.robot_model == HMP_ROBOT_POINT) {
		robot_out = Pose(pose[0], pose[1], pose[2]);
		return;
	}
	double vx = obstacle_pt.x - pose[0], vy = obstacle_pt.y - pose[1];
	double n = std::sqrt(vx * vx + vy * vy);
	robot_out = Pose(pose[0] + (vx / n) * env.robot_radius, pose[1] + (vy / n) * env.robot_radius, pose[2]);
}

// HumapPlanner::enlargeObstacle, src/humap_planner.cpp:681-758
bool enlargeObstacle(const Pose& robot_pt, Pose& obstacle_pt, double extension_distance, double distance_collision_imminent) {
	if (extension_distance <= 0.0) return false;
	V3 dist_init{obstacle_pt.x - robot_pt.x, obstacle_pt.y - robot_pt.y, 0.0};
	if (len3(dist_init) <= distance_collision_imminent) return false;
	double dist_init_dir = std::atan2(dist_init.y, dist_init.x);   // Angle(Vector): not normalised further
	V3 ext{std::cos(dist_init_dir) * extension_distance, std::sin(dist_init_dir) * extension_distance, 0.0};
	Pose hypothesis(obstacle_pt.x - ext.x, obstacle_pt.y - ext.y, obstacle_pt.yaw);
	V3 dist_modded{hypothesis.x - robot_pt.x, hypothesis.y - robot_pt.y, 0.0};
	double dist_angle_diff = std::abs(std::atan2(dist_modded.y, dist_modded.x) - dist_init_dir);
	if (dist_angle_diff <= dtor(1.0)) {
		obstacle_pt = hypothesis;
		return true;
	}
	// fallback stage: the reference re-measures the ORIGINAL closest-point vector here (:744-748), so the test always
	// passes and the obstacle point is placed at the collision distance from the robot point
	hypothesis = Pose(robot_pt.x + distance_collision_imminent * std::cos(dist_init_dir),
	                  robot_pt.y + distance_collision_imminent * std::sin(dist_init_dir), obstacle_pt.yaw);
	dist_modded = {obstacle_pt.x - robot_pt.x, obstacle_pt.y - robot_pt.y, 0.0};
	dist_angle_diff = std::abs(std::atan2(dist_modded.y, dist_modded.x) - dist_init_dir);
	if (dist_angle_diff <= dtor(1.0)) {
		obstacle_pt = hypothesis;
		return true;
	}
	return false;
}

// HumapPlanner::selectRelevant (humap_planner.h:387-427); ties keep the input order (std::sort leaves them unspecified)
std::vector<int> selectRelevant(const std::vector<double>& metric, int max_object_num) {
	std::vector<int> idx(metric.size());
	for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
	if (max_object_num < 0 || metric.size() <= (size_t)max_object_num) return idx;   // (size_t)-1: everything, input order
	if (max_object_num == 0) return {};
	std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return metric[a] < metric[b]; });
	idx.resize((size_t)max_object_num);
	return idx;
}

struct EnvModel {
	std::vector<HmpObstacle> obstacles;   // World::addObstacle call order
	std::vector<int> people, groups;      // people_env_model_, groups_env_model_
};

EnvModel createEnvironmentModel(const HmpEnvParams& env, const double robot_pose[3], const double pose_ref[3], const HmpShape* shapes,
                                int n_shapes, const double* verts, const HmpPerson* people, int n_people, const HmpGroup* groups,
                                int n_groups) {
	EnvModel m;
	// extractNonPeopleObstacles (:760-801)
	std::vector<int> kept;
	for (int i = 0; i < n_shapes; ++i) {
		ShapeView sv{&shapes[i], verts};
		double max_rate = -1.0;
		for (int p = 0; p < n_people; ++p) {
			int within = 0;
			const int np = sv.numPolygonPoints();
			for (int k = 0; k < np; ++k) {
				P2 pt = sv.polygonPoint(k);
				if (std::hypot(pt.x - people[p].x, pt.y - people[p].y) <= env.person_model_radius) within++;
			}
			max_rate = std::max(max_rate, (double)within / np);
		}
		if (n_people > 0 && max_rate >= env.person_containment_rate) continue;
		kept.push_back(i);
	}
	// N closest obstacles / people / groups relative to pose_ (:940-980)
	std::vector<double> metric;
	for (int i : kept) metric.push_back(ShapeView{&shapes[i], verts}.minimumDistance({robot_pose[0], robot_pose[1]}));
	std::vector<int> obs_sel = selectRelevant(metric, env.obstacles_closest_num);
	metric.clear();
	for (int p = 0; p < n_people; ++p) metric.push_back(len3({people[p].x - robot_pose[0], people[p].y - robot_pose[1], 0.0}));
	m.people = selectRelevant(metric, env.people_closest_num);
	metric.clear();
	for (int g = 0; g < n_groups; ++g) metric.push_back(len3({groups[g].x - robot_pose[0], groups[g].y - robot_pose[1], 0.0}));
	m.groups = selectRelevant(metric, env.groups_closest_num);
	auto emit = [&](const Pose& r, const Pose& o, double vx, double vy, double vth, bool force_dynamic) {
		HmpObstacle ob;
		std::memset(&ob, 0, sizeof(ob));
		ob.robot_x = (orc_f64)r.x; ob.robot_y = (orc_f64)r.y; ob.robot_yaw = (orc_f64)r.yaw;
		ob.obj_x = (orc_f64)o.x; ob.obj_y = (orc_f64)o.y; ob.obj_yaw = (orc_f64)o.yaw;
		ob.vx = (orc_f64)vx; ob.vy = (orc_f64)vy; ob.vth = (orc_f64)vth;
		ob.force_dynamic = force_dynamic ? 1 : 0;
		m.obstacles.push_back(ob);
	};
	// :983-1025
	for (int k : obs_sel) {
		const HmpShape& sh = shapes[kept[k]];
		ShapeView sv{&sh, verts};
		Pose r, o;
		calculateClosestPoints(env, pose_ref, sv, r, o);
		enlargeObstacle(r, o, env.obstacle_extension_multiplier * env.robot_radius, 1.05 * env.ttc_collision_distance);
		emit(r, o, sh.vx, sh.vy, 0.0, env.obstacles_force_dynamic != 0);
	}
	// :1027-1049: a person is a CircularObstacle of person_model_radius
	for (int p : m.people) {
		HmpShape circle;
		std::memset(&circle, 0, sizeof(circle));
		circle.type = HMP_SHAPE_CIRCLE;
		circle.x = people[p].x;
		circle.y = people[p].y;
		circle.radius = env.person_model_radius;
		ShapeView sv{&circle, nullptr};
		Pose r, o;
		calculateClosestPoints(env, pose_ref, sv, r, o);
		emit(r, o, people[p].vx, people[p].vy, people[p].vth, env.people_force_dynamic != 0);
	}
	return m;
}

// ------------------------------------------------------------------------------------------------
// Equisampled velocities: base_local_planner::SimpleTrajectoryGenerator + VelocityIterator [RECALLED from upstream
// ros-planning/navigation, parity unpinned] as wired by src/humap_planner.cpp:196-203 and :1317-1361 (first-party).
// The upstream generator works on Eigen::Vector3f: positions, velocities and acceleration limits are FP32 values,
// the expressions between them are evaluated in double (cos/sin/hypot take double, dt is double) and rounded to
// float on every store. Restated with the same store points.
// ------------------------------------------------------------------------------------------------
// base_local_planner/velocity_iterator.h
std::vector<double> velocityIteratorSamples(double vmin, double vmax, int num_samples) {
	std::vector<double> samples;
	if (vmin == vmax) {
		samples.push_back(vmin);
	} else {
		num_samples = std::max(2, num_samples);
		double step_size = (vmax - vmin) / double(std::max(1, (num_samples - 1)));
		double current;
		double next = vmin;
		for (int j = 0; j < num_samples - 1; ++j) {
			current = next;
			next += step_size;
			samples.push_back(current);
			if ((current < 0) && (next > 0)) samples.push_back(0.0);   // a zero between the negative and positive samples
		}
		samples.push_back(vmax);
	}
	return samples;
}

struct EquiGenerator {
	float pos[3], vel[3], acc[3];
	double sim_time = 0, sim_granularity = 0, angular_sim_granularity = 0, sim_period = 0;
	bool continued_acceleration = true, discretize_by_time = true;
	double min_vel_trans = 0, max_vel_trans = 0, min_vel_theta = 0;
	std::vector<std::array<float, 3>> samples;

	// humap_planner.cpp:1317-1361 + SimpleTrajectoryGenerator::initialise
	void initialise(const HmpParams& P, const HmpWorld& w, const HmpEquisampled& eq) {
		const HmpLimits& L = P.limits;
		sim_time = P.general.sim_time;
		sim_granularity = P.general.sim_granularity;
		angular_sim_granularity = P.general.angular_sim_granularity;
		sim_period = P.general.sim_period;
		continued_acceleration = eq.continued_acceleration != 0;   // setParameters(..., use_dwa = !continued, ...)
		const bool use_dwa = !continued_acceleration;
		discretize_by_time = true;                                  // the `true` of humap_planner.cpp:1360
		min_vel_trans = L.min_vel_trans;
		max_vel_trans = L.max_vel_trans;
		min_vel_theta = L.min_vel_theta;
		// :1335-1351 (first-party): keep min_vel_x as high as kinematics allow, but not above max_vel_x
		double maximum_from_min_vel_x = std::max(std::max(eq.min_vel_x, L.min_vel_x),
		                                         std::max(eq.min_vel_x, w.vel_x - L.acc_lim_x * P.general.sim_period));
		double min_vel_x = std::min(maximum_from_min_vel_x, L.max_vel_x);
		double max_vel_x = L.max_vel_x, min_vel_y = L.min_vel_y, max_vel_y = L.max_vel_y;
		double max_vel_th = L.max_vel_theta, min_vel_th = -1.0 * max_vel_th;
		pos[0] = (float)w.robot_x; pos[1] = (float)w.robot_y; pos[2] = (float)w.robot_yaw;   // Pose::getAsEigen2D
		vel[0] = (float)w.vel_x; vel[1] = (float)w.vel_y; vel[2] = (float)w.vel_th;          // Vector::getAsEigen<Vector3f>
		acc[0] = (float)L.acc_lim_x; acc[1] = (float)L.acc_lim_y; acc[2] = (float)L.acc_lim_theta;   // getAccLimits()
		const float goal[2] = {(float)w.goal_x, (float)w.goal_y};
		samples.clear();
		const float vs[3] = {(float)eq.vx_samples, (float)eq.vy_samples, (float)eq.vth_samples};
		if (!(vs[0] * vs[1] * vs[2] > 0)) return;
		float max_vel[3] = {0, 0, 0}, min_vel[3] = {0, 0, 0};
		if (!use_dwa) {
			double dist = std::hypot(goal[0] - pos[0], goal[1] - pos[1]);
			max_vel_x = std::max(std::min(max_vel_x, dist / sim_time), min_vel_x);
			max_vel_y = std::max(std::min(max_vel_y, dist / sim_time), min_vel_y);
			max_vel[0] = (float)std::min(max_vel_x, vel[0] + acc[0] * sim_time);
			max_vel[1] = (float)std::min(max_vel_y, vel[1] + acc[1] * sim_time);
			max_vel[2] = (float)std::min(max_vel_th, vel[2] + acc[2] * sim_time);
			min_vel[0] = (float)std::max(min_vel_x, vel[0] - acc[0] * sim_time);
			min_vel[1] = (float)std::max(min_vel_y, vel[1] - acc[1] * sim_time);
			min_vel[2] = (float)std::max(min_vel_th, vel[2] - acc[2] * sim_time);
		} else {
			max_vel[0] = (float)std::min(max_vel_x, vel[0] + acc[0] * sim_period);
			max_vel[1] = (float)std::min(max_vel_y, vel[1] + acc[1] * sim_period);
			max_vel[2] = (float)std::min(max_vel_th, vel[2] + acc[2] * sim_period);
			min_vel[0] = (float)std::max(min_vel_x, vel[0] - acc[0] * sim_period);
			min_vel[1] = (float)std::max(min_vel_y, vel[1] - acc[1] * sim_period);
			min_vel[2] = (float)std::max(min_vel_th, vel[2] - acc[2] * sim_period);
		}
		auto xs = velocityIteratorSamples(min_vel[0], max_vel[0], (int)vs[0]);
		auto ys = velocityIteratorSamples(min_vel[1], max_vel[1], (int)vs[1]);
		auto ts = velocityIteratorSamples(min_vel[2], max_vel[2], (int)vs[2]);
		for (double vx : xs)
			for (double vy : ys)
				for (double vt : ts) samples.push_back({(float)vx, (float)vy, (float)vt});
	}

	static void computeNewVelocities(const float target[3], const float v[3], const float a[3], double dt, float out[3]) {
		for (int i = 0; i < 3; ++i) {
			if (v[i] < target[i]) out[i] = (float)std::min(double(target[i]), v[i] + a[i] * dt);
			else out[i] = (float)std::max(double(target[i]), v[i] - a[i] * dt);
		}
	}

	// SimpleTrajectoryGenerator::generateTrajectory
	bool generate(size_t k, BlpTrajectory& traj) const {
		const float* target = samples[k].data();
		double vmag = std::hypot(target[0], target[1]);
		const double eps = 1e-4;
		traj.cost = -1.0;
		traj.x.clear();
		traj.y.clear();
		traj.th.clear();
		if ((min_vel_trans >= 0 && vmag + eps < min_vel_trans) && (min_vel_theta >= 0 && std::fabs(target[2]) + eps < min_vel_theta)) return false;
		if (max_vel_trans >= 0 && vmag - eps > max_vel_trans) return false;
		int num_steps;
		if (discretize_by_time) {
			num_steps = (int)std::ceil(sim_time / sim_granularity);
		} else {
			double sim_time_distance = vmag * sim_time;
			double sim_time_angle = std::fabs(target[2]) * sim_time;
			num_steps = (int)std::ceil(std::max(sim_time_distance / sim_granularity, sim_time_angle / angular_sim_granularity));
		}
		if (num_steps == 0) return false;
		double dt = sim_time / num_steps;
		traj.time_delta = dt;
		float loop_vel[3];
		if (continued_acceleration) {
			computeNewVelocities(target, vel, acc, dt, loop_vel);
		} else {
			loop_vel[0] = target[0]; loop_vel[1] = target[1]; loop_vel[2] = target[2];
		}
		traj.xv = loop_vel[0];
		traj.yv = loop_vel[1];
		traj.thetav = loop_vel[2];
		float p[3] = {pos[0], pos[1], pos[2]};
		for (int i = 0; i < num_steps; ++i) {
			traj.x.push_back(p[0]);
			traj.y.push_back(p[1]);
			traj.th.push_back(p[2]);
			if (continued_acceleration) {
				float nv[3];
				computeNewVelocities(target, loop_vel, acc, dt, nv);
				loop_vel[0] = nv[0]; loop_vel[1] = nv[1]; loop_vel[2] = nv[2];
			}
			// computeNewPositions
			float np0 = (float)(p[0] + (loop_vel[0] * std::cos((double)p[2]) + loop_vel[1] * std::cos(M_PI_2 + p[2])) * dt);
			float np1 = (float)(p[1] + (loop_vel[0] * std::sin((double)p[2]) + loop_vel[1] * std::sin(M_PI_2 + p[2])) * dt);
			float np2 = (float)(p[2] + loop_vel[2] * dt);
			p[0] = np0; p[1] = np1; p[2] = np2;
		}
		return true;
	}
};

struct StepForces {
	V3 internal, dynamic, stat, human;
};

// :601-751
StepForces computeForces(const HmpParams& P, FisEngine& fis, const World& world, double dt, const HmpSample& amp) {
	SfmState s;
	// base parameters live in float members (setParameters :293-302), amplified values are truncated to
	// float again (setEquationParameters :73-122 called from :627-637)
	const HmpSfm& c = P.sfm;
	s.relaxation_time = (float)c.relaxation_time;
	s.An = (float)((double)(float)c.an * amp.amp[HMP_AMP_AN]);
	s.Bn = (float)((double)(float)c.bn * amp.amp[HMP_AMP_BN]);
	s.Cn = (float)((double)(float)c.cn * amp.amp[HMP_AMP_CN]);
	s.Ap = (float)((double)(float)c.ap * amp.amp[HMP_AMP_AP]);
	s.Bp = (float)((double)(float)c.bp * amp.amp[HMP_AMP_BP]);
	s.Cp = (float)((double)(float)c.cp * amp.amp[HMP_AMP_CP]);
	s.Aw = (float)((double)(float)c.aw * amp.amp[HMP_AMP_AW]);
	s.Bw = (float)((double)(float)c.bw * amp.amp[HMP_AMP_BW]);
	s.speed_desired = (float)((double)(float)c.speed_desired * amp.amp[HMP_AMP_SPEED]);
	double As = amp.amp[HMP_AMP_AS];  // :641, SURVEY App. A #13
	computeSocialForce(s, c, world, dt);
	V3 human{0, 0, 0};
	if (!c.disable_interaction_forces && !(P.fis.force_factor <= 0.0)) {
		std::vector<FisOutput> outs;
		std::vector<double> speeds, dists, rels;
		double dir_alpha = world.robot.heading_dir;
		for (const Object& o : world.obstacle_dynamic) {
			outs.push_back(fis.processOne(dir_alpha, o.dir_beta, o.rel_loc_angle, o.dist_angle));
			speeds.push_back(o.speed);
			dists.push_back(o.dist);
			rels.push_back(o.rel_loc_angle);
		}
		human = computeBehaviourForce(P.fis, As, world.robot.centroid, world.robot.speed, outs, speeds, dists, rels);
	}
	return {s.force_internal, s.force_dynamic, s.force_static, human};
}

// :292-462; forces_out (optional) receives 8 doubles per step: f_int.xy, f_dyn.xy, f_stat.xy, f_human.xy
bool generateTrajectory(const HmpParams& P, FisEngine& fis, const World& world_model, V3 vel_local, const HmpSample& amp,
                        BlpTrajectory& traj, double* forces_out) {
	const HmpLimits& L = P.limits;
	double speed_linear = std::hypot(vel_local.x, vel_local.y);
	double speed_angular = vel_local.z;
	traj.cost = -1.0;
	traj.x.clear();
	traj.y.clear();
	traj.th.clear();
	int num_steps = computeStepsNumber(P.general, speed_linear, speed_angular);
	double dt = P.general.sim_time / num_steps;
	traj.time_delta = dt;
	World w = world_model;
	for (int i = 0; i < num_steps; ++i) {
		StepForces f = computeForces(P, fis, w, dt, amp);
		if (forces_out) {
			double* o = forces_out + 8 * i;
			o[0] = f.internal.x; o[1] = f.internal.y; o[2] = f.dynamic.x; o[3] = f.dynamic.y;
			o[4] = f.stat.x; o[5] = f.stat.y; o[6] = f.human.x; o[7] = f.human.y;
		}
		V3 total = f.internal + f.dynamic + f.stat + f.human;
		V3 twist = computeTwist(w.robot.centroid, total, P.sfm.mass, L.min_vel_x, L.max_vel_x, L.max_vel_theta,
		                        L.twist_rotation_compensation);
		V3 vel_local_plan = computeVelocityLocal(w.robot.vel, w.robot.centroid);
		twist = adjustTwistWithAccAndGoalLimits(vel_local_plan, L.acc_lim_x, L.acc_lim_y, L.acc_lim_theta, L.min_vel_x,
		                                        L.min_vel_y, -L.max_vel_theta, L.max_vel_x, L.max_vel_y, L.max_vel_theta, dt,
		                                        twist, L.maintain_vel_components_rate != 0, w.robot.goal.dist);
		double sampled_speed_linear = std::hypot(twist.x, twist.y);
		double sampled_speed_angular = twist.z;
		if (!areVelocityLimitsFulfilled(L, sampled_speed_linear, sampled_speed_angular, 1e-4)) {
			return false;
		}
		if (i == 0) {
			traj.xv = twist.x;
			traj.yv = twist.y;
			traj.thetav = twist.z;
		}
		traj.x.push_back(w.robot.centroid.x);
		traj.y.push_back(w.robot.centroid.y);
		traj.th.push_back(w.robot.centroid.yaw);
		V3 twist_glob = computeVelocityGlobal(twist, w.robot.centroid);
		w.predict(twist_glob, dt);
	}
	return true;
}

// Builds World + people/group predictions from the flat scene: HumapPlanner::findBestTrajectory
// (src/humap_planner.cpp:365-370) and humap_planner_ros.cpp:525-554.
void buildScene(PlanState& st, const HmpParams& P, const HmpWorld& hw) {
	Pose pose(hw.robot_x, hw.robot_y, hw.robot_yaw);
	st.vel_local = {hw.vel_x, hw.vel_y, hw.vel_th};
	V3 vel_glob = computeVelocityGlobal(st.vel_local, pose);
	Pose goal_local(hw.goal_local_x, hw.goal_local_y, hw.goal_local_yaw);
	Pose goal(hw.goal_x, hw.goal_y, hw.goal_yaw);
	st.world = World(pose, pose, vel_glob, goal_local, goal);
	for (int i = 0; i < hw.n_obstacles; ++i) {
		const HmpObstacle& o = hw.obstacles[i];
		st.world.addObstacle(Pose(o.robot_x, o.robot_y, o.robot_yaw), Pose(o.obj_x, o.obj_y, o.obj_yaw), V3{o.vx, o.vy, o.vth},
		                     o.force_dynamic != 0);
	}
	unsigned steps = (unsigned)std::ceil(P.general.sim_time / P.general.sim_granularity);
	st.people.clear();
	for (int i = 0; i < hw.n_people; ++i) {
		const HmpPerson& p = hw.people[i];
		PersonPred pp;
		pp.p = p;
		pp.traj = predictObject(Pose(p.x, p.y, p.yaw), V3{p.vx, p.vy, p.vth}, P.general.people_prediction_dt, steps);
		st.people.push_back(pp);
	}
	st.groups.clear();
	for (int i = 0; i < hw.n_groups; ++i) {
		const HmpGroup& g = hw.groups[i];
		GroupPred gp;
		gp.g = g;
		gp.traj = predictObject(Pose(g.x, g.y, g.yaw), V3{0, 0, 0}, 1e-03, steps);
		st.groups.push_back(gp);
	}
}

}  // namespace

// ================================================================================================
// C API
// ================================================================================================
#ifndef HMP_ORACLE_COUNT   /* hmp_oracle_count.cpp compiles everything above with a counting scalar and brings its own entry point */
extern "C" {

int orc_num_candidates(const HmpSampling* sampling, int n_extra) {
	size_t total = 1;
	for (int a = 0; a < HMP_NUM_AMPLIFIERS; ++a) {
		total *= computeAmplifierSamples(sampling->amp_min[a], sampling->amp_max[a], sampling->amp_granularity[a]).size();
	}
	return (int)(total + n_extra);
}

int orc_build_environment(const HmpEnvParams* env, const double robot_pose[3], const double pose_ref[3], const HmpShape* shapes,
                          int32_t n_shapes, const double* vertices_xy, int32_t n_vertices, const HmpPerson* people, int32_t n_people,
                          const HmpGroup* groups, int32_t n_groups, HmpObstacle* obstacles_out, int32_t* n_obstacles_out,
                          int32_t* people_selected, int32_t* n_people_selected, int32_t* groups_selected, int32_t* n_groups_selected) {
	(void)n_vertices;
	EnvModel m = createEnvironmentModel(*env, robot_pose, pose_ref, shapes, n_shapes, vertices_xy, people, n_people, groups, n_groups);
	if ((int)m.obstacles.size() > *n_obstacles_out) return -1;
	std::memcpy(obstacles_out, m.obstacles.data(), m.obstacles.size() * sizeof(HmpObstacle));
	*n_obstacles_out = (int)m.obstacles.size();
	for (size_t i = 0; i < m.people.size(); ++i) people_selected[i] = m.people[i];
	*n_people_selected = (int)m.people.size();
	for (size_t i = 0; i < m.groups.size(); ++i) groups_selected[i] = m.groups[i];
	*n_groups_selected = (int)m.groups.size();
	return 0;
}

// HumapPlanner::computeForceAtPosition (src/humap_planner.cpp:652-678) -> generateTrajectoryWithoutPlanning
// (social_trajectory_generator.cpp:504-527: computeForces with SampleAmplifierSet() and dt = sim_period_)
int orc_force_grid(const HmpParams* P, const HmpEnvParams* env, const HmpWorld* hw, const double* positions_xy, int32_t n_positions,
                   const HmpShape* shapes, int32_t n_shapes, const double* vertices_xy, int32_t n_vertices, double* forces_out) {
	(void)n_vertices;
	const double robot_pose[3] = {hw->robot_x, hw->robot_y, hw->robot_yaw};
	Pose pose_(hw->robot_x, hw->robot_y, hw->robot_yaw);
	V3 vel_glob = computeVelocityGlobal({hw->vel_x, hw->vel_y, hw->vel_th}, pose_);
	Pose goal_local(hw->goal_local_x, hw->goal_local_y, hw->goal_local_yaw), goal(hw->goal_x, hw->goal_y, hw->goal_yaw);
	HmpSample unit;
	for (int a = 0; a < HMP_NUM_AMPLIFIERS; ++a) unit.amp[a] = 1.0;
	FisEngine fis;
	for (int i = 0; i < n_positions; ++i) {
		const double pose_ref[3] = {positions_xy[2 * i], positions_xy[2 * i + 1], hw->robot_yaw};
		Pose pose(pose_ref[0], pose_ref[1], pose_ref[2]);
		World world(pose, pose, vel_glob, goal_local, goal);
		EnvModel m = createEnvironmentModel(*env, robot_pose, pose_ref, shapes, n_shapes, vertices_xy, hw->people, hw->n_people, hw->groups,
		                                    hw->n_groups);
		for (const HmpObstacle& o : m.obstacles)
			world.addObstacle(Pose(o.robot_x, o.robot_y, o.robot_yaw), Pose(o.obj_x, o.obj_y, o.obj_yaw), V3{o.vx, o.vy, o.vth}, o.force_dynamic != 0);
		StepForces f = computeForces(*P, fis, world, P->general.sim_period, unit);
		double* o = forces_out + 8 * (size_t)i;
		o[0] = f.internal.x; o[1] = f.internal.y; o[2] = f.dynamic.x; o[3] = f.dynamic.y;
		o[4] = f.stat.x; o[5] = f.stat.y; o[6] = f.human.x; o[7] = f.human.y;
	}
	return 0;
}

int orc_equisampled_samples(const HmpParams* P, const HmpWorld* w, const HmpEquisampled* eq, double* out, int cap) {
	EquiGenerator g;
	if (eq && eq->enabled) g.initialise(*P, *w, *eq);
	for (size_t i = 0; i < g.samples.size() && (int)i < cap; ++i) {
		out[3 * i] = g.samples[i][0];
		out[3 * i + 1] = g.samples[i][1];
		out[3 * i + 2] = g.samples[i][2];
	}
	return (int)g.samples.size();
}

int orc_num_steps(const HmpParams* P, const HmpWorld* w) {
	return computeStepsNumber(P->general, std::hypot(w->vel_x, w->vel_y), w->vel_th);
}

int orc_samples(const HmpSampling* sampling, const HmpSample* extra, int n_extra, HmpSample* out) {
	auto v = buildSamples(*sampling, extra, n_extra);
	std::memcpy(out, v.data(), v.size() * sizeof(HmpSample));
	return (int)v.size();
}

void orc_mapgrid_compute(const uint8_t* cells, int size_x, int size_y, double origin_x, double origin_y,
                         double resolution, const double* plan_xy, int n_plan, int local_goal, double* target_dist) {
	Costmap cm{cells, size_x, size_y, origin_x, origin_y, resolution};
	mapgridCompute(cm, plan_xy, n_plan, local_goal != 0, target_dist);
}

int orc_plan(const OrcPlanInput* in, OrcPlanOutput* out) {
	PlanState st;
	st.P = in->params;
	st.cm = Costmap{in->cells, in->size_x, in->size_y, in->origin_x, in->origin_y, in->resolution};
	for (int g = 0; g < HMP_NUM_MAPGRIDS; ++g) {
		MapGridCritic& m = st.grids[g];
		m.target_dist = in->target_dist[g];
		m.size_x = in->size_x;
		m.size_y = in->size_y;
		m.xshift = in->params->costs.xshift[g];
		m.yshift = in->params->costs.yshift[g];
		m.stop_on_failure = in->params->costs.stop_on_failure[g] != 0;
		m.n_kernel_size = in->params->costs.neighbour_kernel_size[g];
		m.n_cost_multiplier = in->params->costs.neighbour_cost_multiplier[g];
		m.highest_valid_cost_prev = in->highest_valid_cost_prev[g];
		m.highest_valid_cost = 0.0;
	}
	st.footprint.assign(in->footprint_xy, in->footprint_xy + 2 * in->n_footprint);
	buildScene(st, *in->params, *in->world);
	auto samples = buildSamples(*in->sampling, in->extra, in->n_extra);
	const int n_social = (int)samples.size();
	EquiGenerator equi;
	if (in->equisampled && in->equisampled->enabled) equi.initialise(*in->params, *in->world, *in->equisampled);
	const int C = n_social + (int)equi.samples.size();   // pool order: social generator, then the equisampled one
	const int c_begin = std::max(0, in->cand_begin);
	const int c_end = (in->cand_end <= 0) ? C : std::min(C, in->cand_end);
	FisEngine fis;
	const int T = orc_num_steps(in->params, in->world);

	double best_cost = -1;
	int best_idx = -1;
	BlpTrajectory best_traj;
	double best_raw[HMP_NUM_COSTS];
	int n_generated = 0, n_valid = 0;
	BlpTrajectory traj;
	for (int ci = c_begin; ci < c_end; ++ci) {
		double* forces = (out->forces && ci == out->forces_candidate && ci < n_social) ? out->forces : nullptr;
		bool ok = (ci < n_social) ? generateTrajectory(*in->params, fis, st.world, st.vel_local, samples[ci], traj, forces)
		                          : equi.generate((size_t)(ci - n_social), traj);
		if (out->generated) out->generated[ci] = ok ? 1 : 0;
		if (out->n_poses) out->n_poses[ci] = (int)traj.size();
		if (out->poses) {
			for (size_t i = 0; i < traj.size() && (int)i < T; ++i) {
				double* p = out->poses + ((size_t)ci * T + i) * 3;
				p[0] = traj.x[i];
				p[1] = traj.y[i];
				p[2] = traj.th[i];
			}
		}
		if (out->seeds) {
			out->seeds[3 * ci + 0] = traj.xv;
			out->seeds[3 * ci + 1] = traj.yv;
			out->seeds[3 * ci + 2] = traj.thetav;
		}
		if (!ok) {
			if (out->totals) out->totals[ci] = -1.0;
			if (out->costs) {
				for (int k = 0; k < HMP_NUM_COSTS; ++k) out->costs[(size_t)ci * HMP_NUM_COSTS + k] = std::numeric_limits<double>::quiet_NaN();
			}
			continue;
		}
		n_generated++;
		double raw[HMP_NUM_COSTS];
		double cost = scoreTrajectoryAll(st, traj, best_cost, in->early_exit != 0, raw);
		if (out->totals) out->totals[ci] = cost;
		if (out->costs) std::memcpy(out->costs + (size_t)ci * HMP_NUM_COSTS, raw, sizeof(raw));
		if (cost >= 0) {
			n_valid++;
			if (best_cost < 0 || cost < best_cost) {
				best_cost = cost;
				best_idx = ci;
				best_traj = traj;
				std::memcpy(best_raw, raw, sizeof(raw));
			}
		}
	}
	HmpResult& r = out->result;
	std::memset(&r, 0, sizeof(r));
	r.n_candidates = C;
	r.n_social = n_social;
	r.n_generated = n_generated;
	r.n_valid = n_valid;
	r.best_index = best_idx;
	r.status = best_idx >= 0 ? 0 : 1;
	r.best_total = best_idx >= 0 ? best_cost : -7.0;
	r.time_delta = in->params->general.sim_time / T;
	for (int g = 0; g < HMP_NUM_MAPGRIDS; ++g) r.highest_valid_cost[g] = st.grids[g].highest_valid_cost;
	if (best_idx >= 0) {
		std::memcpy(r.costs, best_raw, sizeof(best_raw));
		r.xv = best_traj.xv;
		r.yv = best_traj.yv;
		r.thetav = best_traj.thetav;
		r.n_poses = (int)best_traj.size();
		for (int a = 0; a < HMP_NUM_AMPLIFIERS; ++a)
			r.amplifiers[a] = (best_idx < n_social) ? samples[best_idx].amp[a] : std::numeric_limits<double>::quiet_NaN();
		if (out->best_poses) {
			for (size_t i = 0; i < best_traj.size(); ++i) {
				out->best_poses[3 * i] = best_traj.x[i];
				out->best_poses[3 * i + 1] = best_traj.y[i];
				out->best_poses[3 * i + 2] = best_traj.th[i];
			}
		}
	}
	return 0;
}

// src/humap_planner.cpp:535-576 over every cell; getFootprintCost(px, py): obstacle_separation_cost_function.cpp:134-145
int orc_cost_cloud(const OrcPlanInput* in, float* cloud6, uint8_t* valid) {
	const HmpCosts& C = in->params->costs;
	Costmap cm{in->cells, in->size_x, in->size_y, in->origin_x, in->origin_y, in->resolution};
	std::vector<double> spec(in->footprint_xy, in->footprint_xy + 2 * in->n_footprint);
	MapGridCritic grids[HMP_NUM_MAPGRIDS];
	for (int g = 0; g < HMP_NUM_MAPGRIDS; ++g) {
		grids[g].target_dist = in->target_dist[g];
		grids[g].size_x = in->size_x;
		grids[g].size_y = in->size_y;
		grids[g].n_kernel_size = C.neighbour_kernel_size[g];
		grids[g].n_cost_multiplier = C.neighbour_cost_multiplier[g];
		grids[g].highest_valid_cost_prev = in->highest_valid_cost_prev[g];
	}
	for (int cy = 0; cy < in->size_y; ++cy) {
		for (int cx = 0; cx < in->size_x; ++cx) {
			const size_t c = (size_t)cy * in->size_x + cx;
			float path_cost = grids[HMP_GRID_PATH].getCellCosts(cx, cy);
			float goal_cost = grids[HMP_GRID_GOAL].getCellCosts(cx, cy);
			double wx, wy;
			cm.mapToWorld(cx, cy, wx, wy);
			float occ_cost = spec.empty() ? -9.0f
			                              : (float)obstacleFootprintCost(cm, wx, wy, 0.0, C.occdist_separation, C.occdist_separation_kernel, spec);
			float align_cost = grids[HMP_GRID_ALIGNMENT].getCellCosts(cx, cy);
			float goal_front_cost = grids[HMP_GRID_GOAL_FRONT].getCellCosts(cx, cy);
			bool unreachable = false;
			const float gv[4] = {path_cost, goal_cost, align_cost, goal_front_cost};
			for (int k = 0; k < 4; ++k)
				unreachable = unreachable || gv[k] == grids[k].obstacleCosts() || gv[k] == grids[k].unreachableCellCosts();
			unreachable = unreachable || occ_cost >= 254 || occ_cost < 0;
			valid[c] = unreachable ? 0 : 1;
			float* o = cloud6 + c * 6;
			if (unreachable) {
				for (int k = 0; k < 6; ++k) o[k] = 0.0f;
				continue;
			}
			path_cost *= C.scale[HMP_COST_PATH];
			goal_cost *= C.scale[HMP_COST_GOAL];
			occ_cost *= C.scale[HMP_COST_OBSTACLE];
			align_cost *= C.scale[HMP_COST_ALIGNMENT];
			goal_front_cost *= C.scale[HMP_COST_GOAL_FRONT];
			o[0] = path_cost + goal_cost + occ_cost + align_cost + goal_front_cost;
			o[1] = path_cost;
			o[2] = goal_cost;
			o[3] = occ_cost;
			o[4] = align_cost;
			o[5] = goal_front_cost;
		}
	}
	return 0;
}

// Scores ONE externally supplied trajectory (poses [n][3], seed twist, dt = sim_time / steps) with all critics
// in the reference order, no early exit: used to check the CUDA critics on the CUDA path's own poses.
int orc_score_trajectory(const OrcPlanInput* in, const double* poses, int n, const double seed[3], double* raw_costs,
                         double* total, double* hv_out) {
	PlanState st;
	st.P = in->params;
	st.cm = Costmap{in->cells, in->size_x, in->size_y, in->origin_x, in->origin_y, in->resolution};
	for (int g = 0; g < HMP_NUM_MAPGRIDS; ++g) {
		MapGridCritic& m = st.grids[g];
		m.target_dist = in->target_dist[g];
		m.size_x = in->size_x;
		m.size_y = in->size_y;
		m.xshift = in->params->costs.xshift[g];
		m.yshift = in->params->costs.yshift[g];
		m.stop_on_failure = in->params->costs.stop_on_failure[g] != 0;
		m.n_kernel_size = in->params->costs.neighbour_kernel_size[g];
		m.n_cost_multiplier = in->params->costs.neighbour_cost_multiplier[g];
		m.highest_valid_cost_prev = in->highest_valid_cost_prev[g];
		m.highest_valid_cost = 0.0;
	}
	st.footprint.assign(in->footprint_xy, in->footprint_xy + 2 * in->n_footprint);
	buildScene(st, *in->params, *in->world);
	BlpTrajectory traj;
	traj.xv = seed[0];
	traj.yv = seed[1];
	traj.thetav = seed[2];
	traj.time_delta = in->params->general.sim_time / orc_num_steps(in->params, in->world);
	for (int i = 0; i < n; ++i) {
		traj.x.push_back(poses[3 * i]);
		traj.y.push_back(poses[3 * i + 1]);
		traj.th.push_back(poses[3 * i + 2]);
	}
	*total = scoreTrajectoryAll(st, traj, -1.0, false, raw_costs);
	if (hv_out) {
		for (int g = 0; g < HMP_NUM_MAPGRIDS; ++g) hv_out[g] = st.grids[g].highest_valid_cost;
	}
	return 0;
}

// ---- known-answer-test hooks (reference test/*.cpp), thin wrappers over the restated functions ----
double orc_wrap(double a) { return wrap(a); }
double orc_yaw_roundtrip(double yaw) { return yaw_roundtrip(yaw); }
double orc_direction(const double v[3]) { return direction({v[0], v[1], v[2]}); }
double orc_len3(const double v[3]) { return len3({v[0], v[1], v[2]}); }
void orc_normalized(const double v[3], double out[3]) {
	V3 n = normalized({v[0], v[1], v[2]});
	out[0] = n.x; out[1] = n.y; out[2] = n.z;
}
void orc_velocity_global(const double vl[3], double yaw, double out[3]) {
	V3 g = computeVelocityGlobal({vl[0], vl[1], vl[2]}, Pose(0, 0, yaw));
	out[0] = g.x; out[1] = g.y; out[2] = g.z;
}
void orc_velocity_local(const double vg[3], double yaw, int holonomic, double out[3]) {
	V3 l = computeVelocityLocal({vg[0], vg[1], vg[2]}, Pose(0, 0, yaw), holonomic != 0);
	out[0] = l.x; out[1] = l.y; out[2] = l.z;
}
void orc_compute_twist(const double pose[3], const double force[3], double mass, double min_vel_x, double max_vel_x,
                       double max_rot_vel, double rot_comp, double out[3]) {
	V3 t = computeTwist(Pose(pose[0], pose[1], pose[2]), {force[0], force[1], force[2]}, mass, min_vel_x, max_vel_x,
	                    max_rot_vel, rot_comp);
	out[0] = t.x; out[1] = t.y; out[2] = t.z;
}
void orc_adjust_twist_acc(const double vel[3], const double acc[3], const double vmin[3], const double vmax[3], double dt,
                          const double cmd[3], int maintain, double out[3]) {
	V3 t = adjustTwistWithAccLimits({vel[0], vel[1], vel[2]}, acc[0], acc[1], acc[2], vmin[0], vmin[1], vmin[2], vmax[0],
	                                vmax[1], vmax[2], dt, {cmd[0], cmd[1], cmd[2]}, maintain != 0);
	out[0] = t.x; out[1] = t.y; out[2] = t.z;
}
void orc_adjust_twist_acc_goal(const double vel[3], const double acc[3], const double vmin[3], const double vmax[3],
                               double dt, const double cmd[3], int maintain, double dist_to_goal, double out[3]) {
	V3 t = adjustTwistWithAccAndGoalLimits({vel[0], vel[1], vel[2]}, acc[0], acc[1], acc[2], vmin[0], vmin[1], vmin[2],
	                                       vmax[0], vmax[1], vmax[2], dt, {cmd[0], cmd[1], cmd[2]}, maintain != 0,
	                                       dist_to_goal);
	out[0] = t.x; out[1] = t.y; out[2] = t.z;
}
void orc_saturate_velocity(const double cmd[3], double max_x, double max_y, double max_trans, double max_theta,
                           double max_back, double out[3]) {
	V3 t = saturateVelocity({cmd[0], cmd[1], cmd[2]}, max_x, max_y, max_trans, max_theta, max_back);
	out[0] = t.x; out[1] = t.y; out[2] = t.z;
}
void orc_next_pose(const double pose[3], const double vel[3], double dt, int base_vel, double out[3]) {
	Pose p(pose[0], pose[1], pose[2]);
	Pose n = base_vel ? computeNextPoseBaseVel(p, {vel[0], vel[1], vel[2]}, dt) : computeNextPose(p, {vel[0], vel[1], vel[2]}, dt);
	out[0] = n.x; out[1] = n.y; out[2] = n.yaw;
}
void orc_internal_force(const double vel[3], const double d[3], double mass, double speed_desired, double relaxation,
                        double out[3]) {
	V3 f = computeInternalForce({vel[0], vel[1], vel[2]}, {d[0], d[1], d[2]}, mass, speed_desired, relaxation);
	out[0] = f.x; out[1] = f.y; out[2] = f.z;
}
double orc_theta_alpha_beta_2011(const double a[3], const double b[3]) {
	return computeThetaAlphaBetaAngle2011({a[0], a[1], a[2]}, {b[0], b[1], b[2]});
}
double orc_theta_alpha_beta_2014(const double n[3], const double d[3]) {
	return computeThetaAlphaBetaAngle2014({n[0], n[1], n[2]}, {d[0], d[1], d[2]});
}
void orc_normal_alpha(double yaw, int description, double out[3]) {
	V3 n = computeNormalAlphaDirection(yaw, description);
	out[0] = n.x; out[1] = n.y; out[2] = n.z;
}
void orc_perpendicular(const double n[3], int rel_loc, int description, double out[3]) {
	V3 p = computePerpendicularToNormal({n[0], n[1], n[2]}, rel_loc, description);
	out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
double orc_relative_speed(const double a[3], const double b[3]) {
	return computeRelativeSpeed({a[0], a[1], a[2]}, {b[0], b[1], b[2]});
}
double orc_factor_fov(double angle, double fov, int gaussian) { return computeFactorFOV(angle, fov, gaussian != 0); }

// blp trajectory -> humap Trajectory velocities (test/test_trajectory.cpp:122-379). Returns #velocities.
int orc_trajectory_velocities(const double* xyth, int n, const double seed[3], double dt, int global, double* vels_out,
                              double* poses_out) {
	BlpTrajectory t;
	t.xv = seed[0]; t.yv = seed[1]; t.thetav = seed[2];
	t.time_delta = dt;
	for (int i = 0; i < n; ++i) {
		t.x.push_back(xyth[3 * i]);
		t.y.push_back(xyth[3 * i + 1]);
		t.th.push_back(xyth[3 * i + 2]);
	}
	Traj tr = makeTrajectory(t, global != 0);
	for (size_t i = 0; i < tr.vels.size(); ++i) {
		vels_out[3 * i] = tr.vels[i].x; vels_out[3 * i + 1] = tr.vels[i].y; vels_out[3 * i + 2] = tr.vels[i].z;
	}
	for (size_t i = 0; i < tr.poses.size(); ++i) {
		poses_out[3 * i] = tr.poses[i].x; poses_out[3 * i + 1] = tr.poses[i].y; poses_out[3 * i + 2] = tr.poses[i].yaw;
	}
	return (int)tr.vels.size();
}

// constant-velocity object prediction (test/test_trajectory.cpp:15-120). Returns #velocities.
int orc_predict_object(const double pose[3], const double vel[3], double dt, int steps, double* poses_out) {
	Traj tr = predictObject(Pose(pose[0], pose[1], pose[2]), {vel[0], vel[1], vel[2]}, dt, steps);
	for (size_t i = 0; i < tr.poses.size(); ++i) {
		poses_out[3 * i] = tr.poses[i].x; poses_out[3 * i + 1] = tr.poses[i].y; poses_out[3 * i + 2] = tr.poses[i].yaw;
	}
	return (int)tr.vels.size();
}

// World + predict (test/test_world_generation.cpp). state_out: per world state
// [cx, cy, cyaw, vx, vy, vz, target_dist, goal_dist, target_angle, n_static, n_dynamic, obj0_x, obj0_y, obj0_dist]
int orc_world_predict_sequence(const HmpWorld* hw, const double* vels, int n_vels, double dt, double* state_out) {
	PlanState st;
	HmpParams P{};
	P.general.sim_time = 1.0;
	P.general.sim_granularity = 1.0;
	P.general.people_prediction_dt = 1.0;
	HmpWorld w2 = *hw;
	w2.n_people = 0;
	w2.n_groups = 0;
	buildScene(st, P, w2);
	// the reference tests build World(pose, vel, ...) with an explicit (global) velocity
	st.world.robot.vel = {hw->vel_x, hw->vel_y, hw->vel_th};
	World w = st.world;
	auto dump = [&](const World& ww, double* o) {
		o[0] = ww.robot.centroid.x; o[1] = ww.robot.centroid.y; o[2] = ww.robot.centroid.yaw;
		o[3] = ww.robot.vel.x; o[4] = ww.robot.vel.y; o[5] = ww.robot.vel.z;
		o[6] = ww.robot.target.dist; o[7] = ww.robot.goal.dist; o[8] = angle_of(ww.robot.target.dist_v);
		o[9] = (double)ww.obstacle_static.size(); o[10] = (double)ww.obstacle_dynamic.size();
		const Object* ob = !ww.obstacle_dynamic.empty() ? &ww.obstacle_dynamic[0] : (!ww.obstacle_static.empty() ? &ww.obstacle_static[0] : nullptr);
		o[11] = ob ? ob->object.x : 0; o[12] = ob ? ob->object.y : 0; o[13] = ob ? ob->dist : 0;
	};
	dump(w, state_out);
	for (int i = 0; i < n_vels; ++i) {
		w.predict({vels[3 * i], vels[3 * i + 1], vels[3 * i + 2]}, dt);
		dump(w, state_out + 14 * (i + 1));
	}
	return n_vels + 1;
}

// fuzzy hooks (test/test_fuzzy_*.cpp)
int orc_trapezoid_update(double intersection_deg, double start, double end, double out[8]) {
	TrapezoidParted tp;
	bool st = trapezoidPartedUpdate(tp, dtor(intersection_deg), start, end);
	for (int k = 0; k < 2; ++k) {
		out[4 * k] = tp.t[k].a; out[4 * k + 1] = tp.t[k].b; out[4 * k + 2] = tp.t[k].c; out[4 * k + 3] = tp.t[k].d;
	}
	return st ? 1 : 0;
}
// side: 1 right, 2 left (RelativeLocation); loc-dependent term update (TrapezoidLocDep::update)
int orc_trapezoid_loc_dep(double intersection_deg, int side, double gamma_start, double gamma_end, double out[8]) {
	if (side == LOC_RIGHT) return orc_trapezoid_update(intersection_deg, wrap(gamma_start), wrap(gamma_end), out);
	return orc_trapezoid_update(intersection_deg, wrap(gamma_end), wrap(gamma_start), out);
}
int orc_trapezoid_loc_indep(double intersection_deg, double length_deg, double gamma_center, double out[8]) {
	double interval = dtor(std::max(length_deg, 1e-03)) / 2.0;
	double gc = wrap(gamma_center);
	return orc_trapezoid_update(intersection_deg, wrap(gc - interval), wrap(gc + interval), out);
}
// out: value, membership, term index; mu_loc[7], mu_dir[5] (after the update) optional
void orc_fis_process(double dir_alpha, double dir_beta, double rel_loc, double dist_angle, double out[3],
                     double* mu_loc, double* mu_dir) {
	FisEngine e;
	FisOutput o = e.processOne(dir_alpha, dir_beta, rel_loc, dist_angle);
	out[0] = o.value; out[1] = o.membership; out[2] = (double)o.term;
	double location = rel_loc > PI ? PI : (rel_loc < -PI ? -PI : rel_loc);
	double dirv = wrap(dir_beta);
	if (mu_loc) for (int i = 0; i < 7; ++i) mu_loc[i] = fl_membership(e.loc_terms[i], location);
	if (mu_dir) for (int i = 0; i < 5; ++i) {
		double a = fl_membership(e.dir_terms[i].t[0], dirv), b = fl_membership(e.dir_terms[i].t[1], dirv);
		mu_dir[i] = a + b - a * b;
	}
}
double orc_behaviour_strength_exp(double action_range, double dist, double speed_agent, double speed_obstacle) {
	return computeBehaviourStrengthExponential(action_range, dist, speed_agent, speed_obstacle);
}
void orc_behaviour_force(const HmpFis* cfg, double As, const double pose[3], double speed_agent, int n,
                         const double* fis_value, const double* fis_membership, const double* speeds,
                         const double* dists, const double* rel_locs, double out[3]) {
	std::vector<FisOutput> outs(n);
	for (int i = 0; i < n; ++i) {
		outs[i].value = fis_value[i];
		outs[i].membership = fis_membership[i];
	}
	V3 f = computeBehaviourForce(*cfg, As, Pose(pose[0], pose[1], pose[2]), speed_agent, outs,
	                             std::vector<double>(speeds, speeds + n), std::vector<double>(dists, dists + n),
	                             std::vector<double>(rel_locs, rel_locs + n));
	out[0] = f.x; out[1] = f.y; out[2] = f.z;
}

// costmap hooks
int orc_world_to_map(const double* wx, const double* wy, int n, int size_x, int size_y, double ox, double oy, double res,
                     int* mx, int* my, int* ok) {
	Costmap cm{nullptr, size_x, size_y, ox, oy, res};
	for (int i = 0; i < n; ++i) {
		unsigned a = 0, b = 0;
		bool r = cm.worldToMap(wx[i], wy[i], a, b);
		ok[i] = r ? 1 : 0;
		mx[i] = r ? (int)a : -1;
		my[i] = r ? (int)b : -1;
	}
	return 0;
}
void orc_footprint_cost(const uint8_t* cells, int size_x, int size_y, double ox, double oy, double res,
                        const double* footprint_xy, int n_fp, const double* xyt, int n, double* cost) {
	Costmap cm{cells, size_x, size_y, ox, oy, res};
	std::vector<double> spec(footprint_xy, footprint_xy + 2 * n_fp);
	for (int i = 0; i < n; ++i) cost[i] = worldModelFootprintCost(cm, xyt[3 * i], xyt[3 * i + 1], xyt[3 * i + 2], spec);
}
void orc_obstacle_cost(const uint8_t* cells, int size_x, int size_y, double ox, double oy, double res,
                       const double* footprint_xy, int n_fp, double separation, int kernel, const double* xyt, int n,
                       double* cost) {
	Costmap cm{cells, size_x, size_y, ox, oy, res};
	std::vector<double> spec(footprint_xy, footprint_xy + 2 * n_fp);
	for (int i = 0; i < n; ++i) {
		cost[i] = obstacleFootprintCost(cm, xyt[3 * i], xyt[3 * i + 1], xyt[3 * i + 2], separation, kernel, spec);
	}
}

double orc_personal_space(const double a[12]) {
	return personalSpaceIntrusion(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11]);
}
double orc_formation_space(const double a[10]) {
	return formationSpaceIntrusion(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]);
}
double orc_heading_disturbance(const double a[15]) {
	return headingDirectionDisturbance(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14]);
}
double orc_passing_speed(double distance, double speed, double min_dist, double max_speed) {
	return passingSpeedDiscomfort(distance, speed, min_dist, max_speed);
}

// ---- hooks for the third-party stand-ins of the oracle/_ref build (oracle/ref_shim/shim_hooks.h): the compiled-in-place
// reference sources call these where the real build would call social_nav_utils / base_local_planner ----------------
double orc_tp_gaussian_angle(double x, double mean, double variance, int normalize) {
	return calculateGaussianAngle(x, mean, variance, normalize != 0);
}
double orc_tp_personal_space(double xp, double yp, double yawp, double cxx, double cxy, double cyx, double cyy, double var_front,
                             double var_rear, double var_side, double xr, double yr) {
	return personalSpaceIntrusion(xp, yp, yawp, cxx, cxy, cyx, cyy, var_front, var_rear, var_side, xr, yr);
}
double orc_tp_formation_space(double xg, double yg, double yawg, double var_x, double var_y, double cxx, double cxy, double cyy,
                              double xr, double yr) {
	return formationSpaceIntrusion(xg, yg, yawg, var_x, var_y, cxx, cxy, cyy, xr, yr);
}
double orc_tp_heading_disturbance(double xp, double yp, double yawp, double cxx, double cxy, double cyy, double xr, double yr,
                                  double yawr, double vxr, double vyr, double person_radius, double fov_person,
                                  double robot_circumradius, double max_speed) {
	return headingDirectionDisturbance(xp, yp, yawp, cxx, cxy, cyy, xr, yr, yawr, vxr, vyr, person_radius, fov_person,
	                                   robot_circumradius, max_speed);
}
double orc_tp_passing_speed(double distance, double speed, double min_dist, double max_speed) {
	return passingSpeedDiscomfort(distance, speed, min_dist, max_speed);
}
double orc_tp_footprint_cost(const uint8_t* cells, int size_x, int size_y, double origin_x, double origin_y, double resolution,
                             double x, double y, double theta, const double* spec_xy, int n_spec) {
	Costmap cm{cells, size_x, size_y, origin_x, origin_y, resolution};
	std::vector<double> spec(spec_xy, spec_xy + 2 * n_spec);
	return worldModelFootprintCost(cm, x, y, theta, spec);
}

}  // extern "C"
#endif  // HMP_ORACLE_COUNT
