"""ORACLE (test infrastructure only): scalar restatement of the reference's real-time dense reward (SURVEY 8(f) N1:
`real_time_rewards` policies - DT baseline, max/min-return modes).  It is the checker of the product's
`dense_reward_kernel` (ctrlsim_b200/csrc/sim.cu, ctrlsim_dense_reward): pinned against the reference's own functions
and against the reference evaluator's DT episode by tests/test_oracle.py, compared with the GPU by
tests/test_gpu_parity.py.  Never imported by the product path.

Restated (reference file:line):
  signed distance to road-edge polylines   utils/data.py:152-290 (_compute_signed_distance_to_polyline(s))
  nearest-vehicle distance                 datasets/rl_waymo/dataset.py:202-236 (normalize=False)
  compute_rewards                          datasets/rl_waymo/dataset.py:239-275
  compute_dense_reward (per step)          evaluators/evaluator.py:106-140
  RTG bookkeeping with real_time_rewards   evaluators/policy_evaluator.py:123-149
Plain Python loops over points / polylines / segments instead of the reference's broadcast arrays.
"""
from __future__ import annotations

import math

import numpy as np

CYCLIC_TOL_M2 = 1.0  # utils/data.py:17


def _sign(x):
    return float(x > 0) - float(x < 0)


def signed_distance_to_polyline(px, py, poly):
    """Negative on the port side (inside), positive on the starboard side of a counter-clockwise boundary."""
    poly = np.asarray(poly, np.float64)
    m = len(poly) - 1  # segments
    cyclic = ((poly[0] - poly[-1]) ** 2).sum() < CYCLIC_TOL_M2
    sx, sy = poly[:-1, 0], poly[:-1, 1]
    ex, ey = poly[1:, 0] - sx, poly[1:, 1] - sy  # start_to_end
    n = np.zeros(m)
    rel = np.zeros(m)
    dist = np.zeros(m)
    for k in range(m):
        ax, ay = px - sx[k], py - sy[k]
        den = ex[k] * ex[k] + ey[k] * ey[k]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.float64(ax * ex[k] + ay * ey[k]) / np.float64(den)
        t = float(np.nan_to_num(t))
        rel[k] = t
        n[k] = _sign(ax * ey[k] - ay * ex[k])
        tc = min(max(t, 0.0), 1.0)
        dist[k] = math.hypot(ax - ex[k] * tc, ay - ey[k] * tc)
    # convexity at every vertex (wrapping the first / last segment direction)
    pex = np.concatenate([ex[-1:], ex, ex[:1]])
    pey = np.concatenate([ey[-1:], ey, ey[:1]])
    convex = pex[:-1] * pey[1:] - pey[:-1] * pex[1:] > 0.0  # [m + 1]
    best = int(np.argmin(dist))
    k = best
    n_prior = (n[-1] if cyclic else n[0]) if k == 0 else n[k - 1]
    n_next = (n[0] if cyclic else n[-1]) if k == m - 1 else n[k + 1]
    if rel[k] < 0.0:
        s = max(n[k], n_prior) if convex[k] else min(n[k], n_prior)
    elif rel[k] < 1.0:
        s = n[k]
    else:
        s = max(n[k], n_next) if convex[k + 1] else min(n[k], n_next)
    return s * dist[best]


def signed_distance_to_polylines(px, py, polylines):
    """The polyline with the smallest |distance| wins (degenerate polylines skipped)."""
    best = None
    for poly in polylines:
        if len(poly) < 2:
            continue
        d = signed_distance_to_polyline(px, py, poly)
        if best is None or abs(d) < abs(best):
            best = d
    return best


def nearest_vehicle_distance(pos, exist):
    """dataset.py:202-236 with normalize=False for ONE step: [n] distances, 0 where undefined."""
    n = len(pos)
    out = np.zeros(n)
    for i in range(n):
        if not exist[i]:
            continue
        best = math.inf
        for j in range(n):
            if j == i or not exist[j]:
                continue
            best = min(best, (pos[i, 0] - pos[j, 0]) ** 2 + (pos[i, 1] - pos[j, 1]) ** 2)
        out[i] = math.sqrt(best) if best < math.inf else 0.0
    return out


def dense_reward_step(w, pos, exist, reward8, road_edge_polylines):
    """evaluators/evaluator.py:106-140 for one step: returns (dense [n, 3] = goal / veh-veh / veh-edge, nearest [n])."""
    n = len(pos)
    dense = np.zeros((n, 3))
    nearest = nearest_vehicle_distance(pos, exist)
    for i in range(n):
        e = float(exist[i])
        r = np.asarray(reward8[i], np.float64) * e
        edge = -signed_distance_to_polylines(pos[i, 0], pos[i, 1], road_edge_polylines) / w.dist_to_road_edge_scaling_factor
        edge *= e
        vv = min(max(nearest[i] * e, 0.0), w.max_veh_veh_distance) / w.max_veh_veh_distance
        if w.remove_shaped_goal:
            goal = r[0] * w.pos_target_achieved_rew_multiplier
        else:
            goal = r[0] * w.pos_target_achieved_rew_multiplier + \
                (min(max(r[3], w.pos_goal_shaped_min), w.pos_goal_shaped_max) - w.pos_goal_shaped_max) * (1 / w.pos_goal_shaped_max)
        if w.remove_shaped_veh_reward:
            veh = -1 * r[6] * w.veh_veh_collision_rew_multiplier
        else:
            veh = vv - r[6] * w.veh_veh_collision_rew_multiplier
        if w.remove_shaped_edge_reward:
            road = -1 * r[7] * w.veh_edge_collision_rew_multiplier
        else:
            road = min(max(abs(edge) * w.dist_to_road_edge_scaling_factor, 0), 5) / 5.0 - r[7] * w.veh_edge_collision_rew_multiplier
        dense[i] = (goal * e, veh * e, road * e)
    return dense, nearest


class RtgTracker:
    """policy_evaluator.py:123-149: un-normalised RTG series of every vehicle under real_time_rewards."""

    def __init__(self, initial_rtgs, evaluated, max_return=False, min_return=False):
        self.rtg = []
        first = np.array(initial_rtgs, np.float64).copy()  # [n, 3]: preproc rtgs[:, 0] with components (0, 3, 4)
        if max_return or min_return:
            first[:] = (10.0, 90.0, 90.0)
        if min_return:
            for v in evaluated:
                first[v] = (0.0, -10.0, -10.0)
        self.rtg.append(first)

    def advance(self, dense_prev):
        self.rtg.append(self.rtg[-1] - np.asarray(dense_prev, np.float64))
        return self.rtg[-1]


def road_edge_polylines(scen_json):
    """evaluators/evaluator.py:143-158 on utils/sim.py:67-73: every point of every 'road_edge' road as the simulator
    holds it - RoadLine::geometry_points() are float32 Vector2D (nocturne/cpp/include/road.h:113), so the file's doubles
    are rounded to float32 before numpy sees them (2e-4 m at Waymo's coordinates of a few thousand metres)."""
    out = []
    for road in scen_json["roads"]:
        geom = road["geometry"]
        if isinstance(geom, dict):
            continue
        if road["type"] == "road_edge":
            out.append(np.array([(np.float32(p["x"]), np.float32(p["y"])) for p in geom], np.float64))
    return out


def initial_rtgs(w, preproc, n):
    """[n, 3] un-normalised RTGs at t = 0 (components 0, 3, 4 of the logged returns): the reverse cumulative sum of
    compute_rewards over the logged episode stored in the *_physics.pkl (datasets/rl_waymo/dataset_ctrl_sim.py:88-92,
    dataset.py:239-275; policy_evaluator.py:125-126)."""
    ag = np.asarray(preproc["ag_data"], np.float64)
    rew = np.asarray(preproc["ag_rewards"], np.float64)
    edge = np.asarray(preproc["veh_edge_dist_rewards"], np.float64)
    vv = np.asarray(preproc["veh_veh_dist_rewards"], np.float64)
    ex = ag[:, :, -1]
    if w.remove_shaped_goal:
        goal = rew[:, :, 0] * w.pos_target_achieved_rew_multiplier
    else:
        goal = rew[:, :, 0] * w.pos_target_achieved_rew_multiplier + \
            (np.clip(rew[:, :, 3], w.pos_goal_shaped_min, w.pos_goal_shaped_max) - w.pos_goal_shaped_max) * (1 / w.pos_goal_shaped_max)
    veh = -1 * rew[:, :, 6] * w.veh_veh_collision_rew_multiplier if w.remove_shaped_veh_reward else \
        vv - rew[:, :, 6] * w.veh_veh_collision_rew_multiplier
    road = -1 * rew[:, :, 7] * w.veh_edge_collision_rew_multiplier if w.remove_shaped_edge_reward else \
        np.clip(np.abs(edge) * w.dist_to_road_edge_scaling_factor, 0, 5) / 5. - rew[:, :, 7] * w.veh_edge_collision_rew_multiplier
    allr = np.stack([goal * ex, veh * ex, road * ex], -1)
    rtgs = np.cumsum(allr[:, ::-1], axis=1)[:, ::-1]
    assert rtgs.shape[0] == n
    return rtgs[:, 0].copy()
