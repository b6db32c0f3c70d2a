// Force-included (-include) into every reference translation unit of the _ref build: the standard headers and logging
// macros that the real ROS / Boost-based headers pull in transitively and the reference relies on implicitly.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <limits>
#include <string>
#include <vector>
#include <ros/console.h>
#include <ros/ros.h>
