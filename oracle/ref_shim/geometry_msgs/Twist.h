// Stand-in: all geometry_msgs structs live in Pose.h
#pragma once
#include <geometry_msgs/Pose.h>
