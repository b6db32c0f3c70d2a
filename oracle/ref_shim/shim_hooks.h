// Hooks the third-party stand-ins call for arithmetic that is NOT first-party reference code. They are exported by
// oracle/hmp_oracle.cpp (libhmp_oracle.so), i.e. the oracle's own statement of the published third-party formulations;
// the _ref build therefore cannot validate them (README.md in this directory).
#pragma once
#include <cstdint>
extern "C" {
double orc_tp_gaussian_angle(double x, double mean, double variance, int normalize);
double orc_tp_personal_space(double xp, double yp, double yawp, double cxx, double cxy, double cyx, double cyy, double var_front,
                             double var_rear, double var_side, double xr, double yr);
double orc_tp_formation_space(double xg, double yg, double yawg, double var_x, double var_y, double cxx, double cxy, double cyy,
                              double xr, double yr);
double orc_tp_heading_disturbance(double xp, double yp, double yawp, double cxx, double cxy, double cyy, double xr, double yr,
                                  double yawr, double vxr, double vyr, double person_radius, double fov_person,
                                  double robot_circumradius, double max_speed);
double orc_tp_passing_speed(double distance, double speed, double min_dist, double max_speed);
double orc_tp_footprint_cost(const uint8_t* cells, int size_x, int size_y, double origin_x, double origin_y, double resolution,
                             double x, double y, double theta, const double* spec_xy, int n_spec);
}
