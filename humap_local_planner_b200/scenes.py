"""Seeded synthetic planning cycles for the BASELINE.json configurations.

Everything here is INPUT generation shared by the tests, bench.py and the CPU baseline: robot state,
people / F-formation groups / obstacle points, the local costmap window (obstacles stamped + inflated
the way costmap_2d's inflation layer does) and the four MapGrid wave-front grids that
MapGridCostFunction::prepare() would hand to the critics (src/map_grid_cost_function.cpp:67-79).
The wave front is the restatement of base_local_planner::MapGrid in numpy; tests/ checks it cell by
cell against the oracle's.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import config
from .capi import (HmpGroup, HmpObstacle, HmpPerson, HmpWorld, Scene)

NO_INFORMATION, LETHAL, INSCRIBED = 255, 254, 253
CORRIDOR_HALF_WIDTH = 1.0


def circular_footprint(radius: float = config.ROBOT_INSCRIBED_RADIUS, n: int = 16) -> np.ndarray:
    """costmap_2d::makeFootprintFromRadius: 16 points on the circle."""
    ang = np.arange(n) * (2.0 * math.pi / n)
    return np.stack([radius * np.cos(ang), radius * np.sin(ang)], axis=1)


def world_to_map(wx, wy, origin_x, origin_y, resolution, size_x, size_y):
    """costmap_2d::Costmap2D::worldToMap."""
    if wx < origin_x or wy < origin_y:
        return None
    mx = int((wx - origin_x) / resolution)
    my = int((wy - origin_y) / resolution)
    if mx < size_x and my < size_y:
        return mx, my
    return None


def inflate(obstacle_mask: np.ndarray, resolution: float, inscribed_radius: float, inflation_radius: float,
            cost_scaling: float = 10.0) -> np.ndarray:
    """costmap_2d::InflationLayer::computeCost on a Euclidean distance transform."""
    from scipy import ndimage
    dist = ndimage.distance_transform_edt(~obstacle_mask) * resolution
    cells = np.zeros(obstacle_mask.shape, dtype=np.uint8)
    ring = (dist > inscribed_radius) & (dist <= inflation_radius)
    cells[ring] = ((INSCRIBED - 1) * np.exp(-cost_scaling * (dist[ring] - inscribed_radius))).astype(np.uint8)
    cells[(dist > 0) & (dist <= inscribed_radius)] = INSCRIBED
    cells[obstacle_mask] = LETHAL
    return cells


def adjust_plan_resolution(plan_xy: np.ndarray, resolution: float) -> np.ndarray:
    """base_local_planner::MapGrid::adjustPlanResolution."""
    out = [plan_xy[0]]
    last = plan_xy[0]
    for p in plan_xy[1:]:
        sq = float((p[0] - last[0]) ** 2 + (p[1] - last[1]) ** 2)
        if sq > resolution * resolution:
            steps = math.ceil(math.sqrt(sq) / resolution)
            d = (p - last) / steps
            for j in range(1, steps):
                out.append(last + j * d)
        out.append(p)
        last = p
    return np.array(out)


def mapgrid_wavefront(cells: np.ndarray, origin_x: float, origin_y: float, resolution: float, plan_xy: np.ndarray,
                      local_goal: bool) -> np.ndarray:
    """base_local_planner::MapGrid::{setTargetCells, setLocalGoal} + computeTargetDistance, vectorised.

    4-connected wave front from the seed cells; cells whose cost is LETHAL / INSCRIBED / NO_INFORMATION and
    that the front touches get obstacleCosts() = size_x*size_y; untouched cells keep
    unreachableCellCosts() = size_x*size_y + 1.
    """
    size_y, size_x = cells.shape
    obstacle_costs = float(size_x * size_y)
    unreachable = obstacle_costs + 1.0
    dist = np.full(cells.shape, unreachable, dtype=np.float64)
    if plan_xy is None or len(plan_xy) == 0:
        return dist
    pts = adjust_plan_resolution(np.asarray(plan_xy, dtype=np.float64), resolution)
    seeds = []
    started = False
    for p in pts:
        m = world_to_map(p[0], p[1], origin_x, origin_y, resolution, size_x, size_y)
        if m is not None and cells[m[1], m[0]] != NO_INFORMATION:
            if local_goal:
                seeds = [m]
            else:
                seeds.append(m)
            started = True
        elif started:
            break
    if not started:
        return dist
    mark = np.zeros(cells.shape, dtype=bool)
    frontier = np.zeros(cells.shape, dtype=bool)
    for mx, my in seeds:
        dist[my, mx] = 0.0
        mark[my, mx] = True
        frontier[my, mx] = True
    blocked = (cells == LETHAL) | (cells == INSCRIBED) | (cells == NO_INFORMATION)
    level = 0.0
    while frontier.any():
        level += 1.0
        nb = np.zeros_like(frontier)
        nb[:, 1:] |= frontier[:, :-1]
        nb[:, :-1] |= frontier[:, 1:]
        nb[1:, :] |= frontier[:-1, :]
        nb[:-1, :] |= frontier[1:, :]
        nb &= ~mark
        mark |= nb
        hit = nb & blocked
        dist[hit] = obstacle_costs
        frontier = nb & ~blocked
        dist[frontier] = level
    return dist


@dataclass
class CycleConfig:
    """One BASELINE.json configuration."""
    name: str
    n_people: int
    n_groups: int
    n_obstacles: int
    sim_time: float
    sim_granularity: float
    sampling: dict
    size: int = 200
    resolution: float = 0.05


CONFIGS = {
    # Reference CPU planning cycle: 4 people incl. one 2-person F-formation, 30 obstacle points, default sampling, 3.5 s @ 0.1 s
    "cfg0": CycleConfig("cfg0", 4, 1, 30, 3.5, 0.1, config.SAMPLING_CFG_DEFAULT),
    # Dense sampling: 16k candidates, 10 people, 2 groups, 200x200 @ 0.05 m
    "cfg1": CycleConfig("cfg1", 10, 2, 30, 3.5, 0.1, config.SAMPLING_16K),
    # Crowd stress: 64k candidates x 50 people x 8 groups x 500 obstacle points, 5 s horizon
    "cfg2": CycleConfig("cfg2", 50, 8, 500, 5.0, 0.1, config.SAMPLING_64K),
    # Batched scenes: cfg1 geometry, 4k candidates per scene
    "cfg3": CycleConfig("cfg3", 10, 2, 30, 3.5, 0.1, config.SAMPLING_4K),
}


def make_scene(cfg: CycleConfig, seed: int, robot_xy=(0.0, 0.0), robot_yaw: float = 0.0,
               base_vel=(0.3, 0.0, 0.0), with_grids: bool = True) -> Scene:
    """Seeded synthetic World + costmap + MapGrids + footprint for one planning cycle. with_grids=False skips the host wave
    fronts (Scene.grids is empty; Scene.plans holds what hmp_compute_mapgrid[_batch] needs to compute them on the device)."""
    rng = np.random.default_rng(seed)
    res, n = cfg.resolution, cfg.size
    rx, ry = robot_xy
    half = 0.5 * n * res
    origin_x, origin_y = rx - half, ry - half
    r_robot = config.ROBOT_INSCRIBED_RADIUS
    r_person = config.PERSON_MODEL_RADIUS

    # --- static obstacle points: annulus 0.6 .. 4.5 m, kept clear of the straight corridor right in front
    # (|lateral| >= CORRIDOR_HALF_WIDTH for -1 m <= longitudinal <= 5 m) so that a useful share of the candidates
    # is collision-free even with 500 points
    oxl, oyl = [], []
    while len(oxl) < cfg.n_obstacles:
        a = rng.uniform(-math.pi, math.pi)
        r = rng.uniform(0.6, 4.5)
        lx, ly = r * math.cos(a), r * math.sin(a)   # robot frame
        if -1.0 <= lx <= 5.0 and abs(ly) < CORRIDOR_HALF_WIDTH:
            continue
        oxl.append(lx)
        oyl.append(ly)
    oxl, oyl = np.array(oxl), np.array(oyl)
    cy0, sy0 = math.cos(robot_yaw), math.sin(robot_yaw)
    ox, oy = rx + oxl * cy0 - oyl * sy0, ry + oxl * sy0 + oyl * cy0
    mask = np.zeros((n, n), dtype=bool)
    for x, y in zip(ox, oy):
        m = world_to_map(x, y, origin_x, origin_y, res, n, n)
        if m is not None:
            mask[m[1], m[0]] = True
    cells = inflate(mask, res, r_robot, 0.55)
    if seed % 2 == 1:  # a 2-cell NO_INFORMATION border in half of the seeds
        cells[:2, :] = cells[-2:, :] = NO_INFORMATION
        cells[:, :2] = cells[:, -2:] = NO_INFORMATION

    obstacles: List[HmpObstacle] = []
    for x, y in zip(ox, oy):
        d = math.hypot(x - rx, y - ry)
        ux, uy = (x - rx) / d, (y - ry) / d
        o = HmpObstacle()
        o.robot_x, o.robot_y, o.robot_yaw = rx + r_robot * ux, ry + r_robot * uy, robot_yaw
        o.obj_x, o.obj_y, o.obj_yaw = x, y, 0.0
        o.vx = o.vy = o.vth = 0.0
        o.force_dynamic = 0
        obstacles.append(o)

    # --- people: annulus 1 .. 5 m; the first 2*G are paired into (almost standing) F-formations
    people: List[HmpPerson] = []
    groups: List[HmpGroup] = []
    P, G = cfg.n_people, cfg.n_groups

    def person(x, y, speed, heading):
        p = HmpPerson()
        p.x, p.y = x, y
        p.vx, p.vy, p.vth = speed * math.cos(heading), speed * math.sin(heading), 0.0
        p.yaw = math.atan2(p.vy, p.vx)  # people_msgs_utils::Person: orientation = atan2(vy, vx), test_trajectory.cpp:25
        p.cov_xx = p.cov_yy = 0.05 ** 2
        p.cov_xy = p.cov_yx = 0.0
        return p

    for gi in range(G):
        a = rng.uniform(-math.pi, math.pi)
        r = rng.uniform(1.5, 4.5)
        cx, cy = rx + r * math.cos(a), ry + r * math.sin(a)
        axis = rng.uniform(-math.pi, math.pi)
        sep = rng.uniform(0.8, 1.2)
        p1 = (cx + 0.5 * sep * math.cos(axis), cy + 0.5 * sep * math.sin(axis))
        p2 = (cx - 0.5 * sep * math.cos(axis), cy - 0.5 * sep * math.sin(axis))
        # members face each other and barely move: exercises the static<->dynamic re-classification (App. A #5)
        people.append(person(p1[0], p1[1], rng.uniform(0.0, 0.03), axis + math.pi))
        people.append(person(p2[0], p2[1], rng.uniform(0.0, 0.03), axis))
        g = HmpGroup()
        g.x, g.y, g.yaw = cx, cy, axis
        g.span_x, g.span_y = rng.uniform(0.8, 2.0), rng.uniform(0.8, 2.0)
        g.cov_xx = g.cov_yy = 0.05 ** 2
        g.cov_xy = 0.0
        groups.append(g)
    while len(people) < P:
        a = rng.uniform(-math.pi, math.pi)
        r = rng.uniform(1.0, 5.0)
        people.append(person(rx + r * math.cos(a), ry + r * math.sin(a), rng.uniform(0.0, 1.5),
                             rng.uniform(-math.pi, math.pi)))
    people = people[:P]

    # people enter the World as circles of radius person_model_radius, dynamic formulation forced
    # (humap_planner.cpp:1016-1036, human_force_formulation_dynamic = true)
    for p in people:
        d = math.hypot(p.x - rx, p.y - ry)
        ux, uy = (p.x - rx) / d, (p.y - ry) / d
        o = HmpObstacle()
        o.robot_x, o.robot_y, o.robot_yaw = rx + r_robot * ux, ry + r_robot * uy, robot_yaw
        o.obj_x, o.obj_y, o.obj_yaw = p.x - r_person * ux, p.y - r_person * uy, 0.0
        o.vx, o.vy, o.vth = p.vx, p.vy, p.vth
        o.force_dynamic = 1
        obstacles.append(o)

    # --- goals and plan (straight line along the robot's heading)
    hx, hy = math.cos(robot_yaw), math.sin(robot_yaw)
    goal = (rx + 8.0 * hx, ry + 8.0 * hy)
    goal_local = (rx + 4.0 * hx, ry + 4.0 * hy)
    fwd = 0.325
    plan = np.array([[rx + s * hx, ry + s * hy] for s in np.arange(0.0, 4.0 + 1e-9, 0.1)])
    front_plan = np.array([[rx + s * hx, ry + s * hy] for s in np.arange(0.0, fwd, 0.1)] + [[rx + fwd * hx, ry + fwd * hy]])
    grids = [
        mapgrid_wavefront(cells, origin_x, origin_y, res, plan, False),        # path_costs_
        mapgrid_wavefront(cells, origin_x, origin_y, res, plan, True),         # goal_costs_ (local goal)
        mapgrid_wavefront(cells, origin_x, origin_y, res, plan, False),        # alignment_costs_
        mapgrid_wavefront(cells, origin_x, origin_y, res, front_plan, True),   # goal_front_costs_ (local goal)
    ] if with_grids else []

    w = HmpWorld()
    w.robot_x, w.robot_y, w.robot_yaw = rx, ry, robot_yaw
    w.vel_x, w.vel_y, w.vel_th = base_vel
    w.goal_local_x, w.goal_local_y, w.goal_local_yaw = goal_local[0], goal_local[1], robot_yaw
    w.goal_x, w.goal_y, w.goal_yaw = goal[0], goal[1], robot_yaw
    obs_arr = (HmpObstacle * max(1, len(obstacles)))(*obstacles)
    ppl_arr = (HmpPerson * max(1, len(people)))(*people)
    grp_arr = (HmpGroup * max(1, len(groups)))(*groups)
    w.obstacles, w.people, w.groups = obs_arr, ppl_arr, grp_arr
    w.n_obstacles, w.n_people, w.n_groups = len(obstacles), len(people), len(groups)
    hv_prev = (80.0, 80.0, 80.0, 80.0)
    sc = Scene(w, obs_arr, ppl_arr, grp_arr, cells, origin_x, origin_y, res, grids, circular_footprint(), hv_prev)
    # the target poses each MapGridCostFunction received (setTargetPoses) and its is_local_goal_function_ flag
    sc.plans = [(plan, False), (plan, True), (plan, False), (front_plan, True)]
    return sc


def make_params(cfg: CycleConfig, fis: bool = True):
    return config.default_params(cfg.resolution, cfg.sim_time, cfg.sim_granularity, fis)


def make_sampling(cfg: CycleConfig):
    return config.make_sampling(cfg.sampling)


def make_shapes(seed: int, n: int = 40, people=None, robot_xy=(0.0, 0.0)):
    """Synthetic costmap_converter output: a mix of point, circle, line and polygon obstacles around the robot, a few of
    them sitting on people (legs seen by the laser) so that extractNonPeopleObstacles has something to remove.
    Returns (ctypes array of HmpShape, vertex pool [m][2])."""
    from .capi import HmpShape, SHAPE_POINT, SHAPE_CIRCLE, SHAPE_LINE, SHAPE_POLYGON
    rng = np.random.default_rng(1000 + seed)
    shapes, verts = [], []
    rx, ry = robot_xy
    for i in range(n):
        s = HmpShape()
        a, r = rng.uniform(-math.pi, math.pi), rng.uniform(0.7, 4.5)
        cx, cy = rx + r * math.cos(a), ry + r * math.sin(a)
        kind = i % 4
        s.vx, s.vy = (rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3)) if i % 7 == 0 else (0.0, 0.0)
        if kind == 0:
            s.type, s.x, s.y = SHAPE_POINT, cx, cy
        elif kind == 1:
            s.type, s.x, s.y, s.radius = SHAPE_CIRCLE, cx, cy, rng.uniform(0.05, 0.3)
        elif kind == 2:
            th, ln = rng.uniform(-math.pi, math.pi), rng.uniform(0.2, 1.5)
            s.type, s.x, s.y = SHAPE_LINE, cx, cy
            s.x2, s.y2 = cx + ln * math.cos(th), cy + ln * math.sin(th)
        else:
            k = int(rng.integers(3, 7))
            ang = np.sort(rng.uniform(-math.pi, math.pi, k))
            rad = rng.uniform(0.1, 0.5, k)
            s.type, s.first_vertex, s.n_vertices = SHAPE_POLYGON, len(verts), k
            for t, q in zip(ang, rad):
                verts.append((cx + q * math.cos(t), cy + q * math.sin(t)))
        shapes.append(s)
    if people is not None:
        for p in list(people)[:3]:   # leg-like obstacles inside the person model radius
            s = HmpShape()
            s.type, s.first_vertex, s.n_vertices = SHAPE_POLYGON, len(verts), 3
            for t in (0.0, 2.1, 4.2):
                verts.append((p.x + 0.1 * math.cos(t), p.y + 0.1 * math.sin(t)))
            shapes.append(s)
            s2 = HmpShape()
            s2.type, s2.x, s2.y = SHAPE_POINT, p.x + 0.05, p.y - 0.05
            shapes.append(s2)
    arr = (HmpShape * len(shapes))(*shapes)
    return arr, np.array(verts, dtype=np.float64).reshape(-1, 2)


def make_env_params(closest=(-1, -1, -1), robot_model: int = 1):
    from .capi import HmpEnvParams
    e = HmpEnvParams()
    e.robot_model = robot_model
    e.obstacles_closest_num, e.people_closest_num, e.groups_closest_num = closest
    e.robot_radius = config.ROBOT_INSCRIBED_RADIUS
    e.person_model_radius = config.PERSON_MODEL_RADIUS
    e.obstacle_extension_multiplier = 1.0
    e.ttc_collision_distance = 0.05
    e.person_containment_rate = 0.667
    e.obstacles_force_dynamic = 0
    e.people_force_dynamic = 1
    # geometry of the non-circular footprint models (robot frame): two circles, a line, a 0.7 m x 0.5 m box with a nose
    e.two_circles[:] = [0.2, 0.25, -0.2, 0.3]
    e.line_xy[:] = [-0.3, 0.0, 0.3, 0.0]
    poly = [(0.35, 0.15), (0.45, 0.0), (0.35, -0.15), (-0.35, -0.25), (-0.35, 0.25)]
    e.n_polygon = len(poly)
    for i, (px, py) in enumerate(poly):
        e.polygon_xy[2 * i], e.polygon_xy[2 * i + 1] = px, py
    return e
