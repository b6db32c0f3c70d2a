// Stand-in for costmap_2d/cost_values.h (published constants).
#pragma once
namespace costmap_2d {
static const unsigned char NO_INFORMATION = 255;
static const unsigned char LETHAL_OBSTACLE = 254;
static const unsigned char INSCRIBED_INFLATED_OBSTACLE = 253;
static const unsigned char FREE_SPACE = 0;
}  // namespace costmap_2d
