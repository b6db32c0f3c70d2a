// Stand-in for costmap_2d/costmap_2d.h: a non-owning view of a row-major uint8 grid with the published
// worldToMap / mapToWorld / getCost semantics (third-party; restated, not validated by the _ref build).
#pragma once
#include <costmap_2d/cost_values.h>
#include <geometry_msgs/Point.h>
#include <vector>
namespace costmap_2d {
class Costmap2D {
public:
	Costmap2D() : cells_(nullptr), size_x_(0), size_y_(0), origin_x_(0), origin_y_(0), resolution_(1) {}
	Costmap2D(const unsigned char* cells, unsigned int size_x, unsigned int size_y, double resolution, double origin_x,
	          double origin_y)
	    : cells_(cells), size_x_(size_x), size_y_(size_y), origin_x_(origin_x), origin_y_(origin_y), resolution_(resolution) {}
	unsigned char getCost(unsigned int mx, unsigned int my) const { return cells_[(size_t)my * size_x_ + mx]; }
	bool worldToMap(double wx, double wy, unsigned int& mx, unsigned int& my) const {
		if (wx < origin_x_ || wy < origin_y_) return false;
		mx = (int)((wx - origin_x_) / resolution_);
		my = (int)((wy - origin_y_) / resolution_);
		return mx < size_x_ && my < size_y_;
	}
	void mapToWorld(unsigned int mx, unsigned int my, double& wx, double& wy) const {
		wx = origin_x_ + (mx + 0.5) * resolution_;
		wy = origin_y_ + (my + 0.5) * resolution_;
	}
	unsigned int getSizeInCellsX() const { return size_x_; }
	unsigned int getSizeInCellsY() const { return size_y_; }
	double getOriginX() const { return origin_x_; }
	double getOriginY() const { return origin_y_; }
	double getResolution() const { return resolution_; }
	const unsigned char* getCharMap() const { return cells_; }
private:
	const unsigned char* cells_;
	unsigned int size_x_, size_y_;
	double origin_x_, origin_y_, resolution_;
};
}  // namespace costmap_2d
