// Stand-in for base_local_planner/map_grid.h. The wave front itself (setTargetCells / setLocalGoal ->
// computeTargetDistance) is third-party: this MapGrid takes its target_dist values from a fill hook installed by
// oracle/ref_driver.cpp, which hands over the very grids the test passes to the oracle and to the CUDA path.
#pragma once
#include <costmap_2d/costmap_2d.h>
#include <geometry_msgs/PoseStamped.h>
#include <functional>
#include <vector>
namespace base_local_planner {
class MapCell {
public:
	unsigned int cx = 0, cy = 0;
	double target_dist = 0.0;
	bool target_mark = false;
	bool within_robot = false;
};
class MapGrid {
public:
	typedef std::function<void(MapGrid&, const costmap_2d::Costmap2D&, const std::vector<geometry_msgs::PoseStamped>&, bool)> FillHook;
	MapGrid() : size_x_(0), size_y_(0) {}
	MapGrid(unsigned int size_x, unsigned int size_y) : size_x_(size_x), size_y_(size_y) { map_.resize((size_t)size_x * size_y); }
	// upstream: map_[size_x_ * y + x] in unsigned int arithmetic. The reference's neighbour heuristic
	// (src/map_grid_cost_function.cpp:95-126) can form indices outside the array, which is undefined
	// behaviour upstream; here such an index yields a scratch cell (target_dist 0) instead of touching foreign memory.
	MapCell& operator()(unsigned int x, unsigned int y) {
		unsigned int idx = size_x_ * y + x;
		if (idx >= map_.size()) {
			scratch_ = MapCell();
			return scratch_;
		}
		return map_[idx];
	}
	MapCell& getCell(unsigned int x, unsigned int y) { return (*this)(x, y); }
	// published: obstacles map_.size(), unreachable cells map_.size() + 1
	double obstacleCosts() { return (double)map_.size(); }
	double unreachableCellCosts() { return (double)map_.size() + 1; }
	void sizeCheck(unsigned int size_x, unsigned int size_y) {
		if (map_.size() != (size_t)size_x * size_y) map_.resize((size_t)size_x * size_y);
		size_x_ = size_x;
		size_y_ = size_y;
	}
	void resetPathDist() {
		for (auto& c : map_) {
			c.target_dist = unreachableCellCosts();
			c.target_mark = false;
			c.within_robot = false;
		}
	}
	void setTargetCells(const costmap_2d::Costmap2D& costmap, const std::vector<geometry_msgs::PoseStamped>& global_plan) {
		sizeCheck(costmap.getSizeInCellsX(), costmap.getSizeInCellsY());
		if (fillHook()) fillHook()(*this, costmap, global_plan, false);
	}
	void setLocalGoal(const costmap_2d::Costmap2D& costmap, const std::vector<geometry_msgs::PoseStamped>& global_plan) {
		sizeCheck(costmap.getSizeInCellsX(), costmap.getSizeInCellsY());
		if (fillHook()) fillHook()(*this, costmap, global_plan, true);
	}
	static FillHook& fillHook() {
		static thread_local FillHook hook;
		return hook;
	}
	unsigned int size_x_, size_y_;
private:
	std::vector<MapCell> map_;
	MapCell scratch_;
};
}  // namespace base_local_planner
