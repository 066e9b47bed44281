"""Host-side scenario loading: Nocturne scenario JSON (+ preprocessed road pkl) -> flat arrays for the GPU batch.

Replaces, for the evaluation path, what the reference spreads over
  Scenario::LoadObjects / LoadRoads          nocturne/cpp/src/scenario.cc:893-1057  (objects valid at t=0 only,
                                             heading deg -> rad in float + NormalizeAngle, speed = |velocity|,
                                             road_edge polylines -> collision segments, is_moving rule)
  get_ground_truth_states                    utils/sim.py:20-65   (expert replay = the JSON tracks themselves)
  Evaluator.initialize_goal_dict             evaluators/evaluator.py:60-84
  PolicyEvaluator.evaluate_policy (setup)    evaluators/policy_evaluator.py:440-505 (moving vehicles, random.sample of
                                             the evaluated set, descending-GT-length focal order)
Values are kept in the precision the reference holds them in (float32 simulator quantities, float64 python-side).
"""
from __future__ import annotations

import math

from operator import itemgetter

import numpy as np


def _normalize_angle_f32(deg: float) -> np.float32:
    """geometry_utils.h:41-58 instantiated with T = float."""
    rad = np.float32(float(np.float32(deg)) / 180.0 * math.pi)
    ret = np.float32(math.fmod(float(rad), 2.0 * math.pi))
    r = float(ret)
    if r > math.pi:
        r = r - 2.0 * math.pi
    elif r < -math.pi:
        r = r + 2.0 * math.pi
    return np.float32(r)


_GET_X, _GET_Y = itemgetter("x"), itemgetter("y")


def _xy_array(points) -> np.ndarray:
    """[{'x': .., 'y': ..}, ...] -> [n, 2] float32.  Two flat lists convert about twice as fast as a list of pairs
    (the parser is ~6 ms of Python per 64-vehicle scene, the only host work of an evaluation that scales with it)."""
    out = np.empty((len(points), 2), np.float32)
    out[:, 0] = list(map(_GET_X, points))
    out[:, 1] = list(map(_GET_Y, points))
    return out


def parse_scenario(scen: dict, steps: int = 90, moving_threshold: float = 0.2, speed_threshold: float = 0.05,
                   vectorize: bool = True):
    """``vectorize=False`` forces the per-vehicle loop (what scenes with tracks of different lengths take); both paths
    give bit-identical arrays (tests/test_oracle.py::test_parser_paths_are_bit_identical)."""
    objs = [o for o in scen["objects"] if bool(o["valid"][0]) and o["type"] == "vehicle"]
    n, T1 = len(objs), steps + 1
    gt = np.zeros((n, T1, 4), np.float32)
    gt_valid = np.zeros((n, T1), np.uint8)
    size = np.zeros((n, 2), np.float32)
    target = np.zeros((n, 4), np.float32)
    moving = np.zeros(n, bool)
    two_pi = 2.0 * math.pi
    lengths = {len(o["position"]) for o in objs}
    if vectorize and len(lengths) == 1 and min(lengths) >= T1:
        # every track has the same length (the Waymo-derived files: 91 states): all vehicles at once, the same
        # element-wise float32 / float64 arithmetic as the per-vehicle loop below (bit-identical, ~3x faster)
        L = lengths.pop()
        pos, vel = np.empty((n, L, 2), np.float32), np.empty((n, L, 2), np.float32)
        pos[:, :, 0] = [list(map(_GET_X, o["position"])) for o in objs]
        pos[:, :, 1] = [list(map(_GET_Y, o["position"])) for o in objs]
        vel[:, :, 0] = [list(map(_GET_X, o["velocity"])) for o in objs]
        vel[:, :, 1] = [list(map(_GET_Y, o["velocity"])) for o in objs]
        size[:] = [(o["length"], o["width"]) for o in objs]
        gps = [o.get("goalPosition", {"x": 0.0, "y": 0.0}) for o in objs]
        target[:, 0], target[:, 1] = list(map(_GET_X, gps)), list(map(_GET_Y, gps))
        deg = np.asarray([o["heading"] for o in objs], np.float32).reshape(n, L)
        rad = (deg.astype(np.float64) / 180.0 * math.pi).astype(np.float32)
        r = np.fmod(rad.astype(np.float64), two_pi).astype(np.float32).astype(np.float64)
        h = np.where(r > math.pi, r - two_pi, np.where(r < -math.pi, r + two_pi, r)).astype(np.float32)
        sp = np.sqrt(vel[:, :, 0] * vel[:, :, 0] + vel[:, :, 1] * vel[:, :, 1])
        gt[:, :, 0], gt[:, :, 1], gt[:, :, 2], gt[:, :, 3] = pos[:, :T1, 0], pos[:, :T1, 1], h[:, :T1], sp[:, :T1]
        gt_valid[:] = pos[:, :T1, 0] != np.float32(-10000.0)  # utils/sim.py:28 existence rule
        valid = np.asarray([o["valid"][:L] for o in objs], bool).reshape(n, L)
        has = valid.any(axis=1)
        last = L - 1 - np.argmax(valid[:, ::-1], axis=1)
        rows = np.arange(n)
        target[:, 2] = np.where(has, h[rows, last], np.float32(0))
        target[:, 3] = np.where(has, sp[rows, last], np.float32(0))
        dx, dy = pos[:, :, 0] - target[:, None, 0], pos[:, :, 1] - target[:, None, 1]
        dist = np.sqrt(dx * dx + dy * dy)
        moving[:] = (((sp > np.float32(speed_threshold)) | (dist > np.float32(moving_threshold))) & valid).any(axis=1)
        objs_loop = []
    else:
        objs_loop = objs
    for i, o in enumerate(objs_loop):
        L = len(o["position"])
        if L < T1:
            raise ValueError(f"object {i}: {L} states, the evaluator needs {T1}")
        size[i] = (o["length"], o["width"])
        gp = o.get("goalPosition", {"x": 0.0, "y": 0.0})
        target[i, :2] = (gp["x"], gp["y"])
        # whole track at once, in the arithmetic of the scalar rules above (float32 simulator values; the angle goes
        # float32 -> double -> float32 twice exactly like geometry_utils.h:41-58 with T = float)
        pos, vel = _xy_array(o["position"]), _xy_array(o["velocity"])
        deg = np.asarray(o["heading"], np.float32)
        rad = (deg.astype(np.float64) / 180.0 * math.pi).astype(np.float32)
        r = np.fmod(rad.astype(np.float64), two_pi).astype(np.float32).astype(np.float64)
        h = np.where(r > math.pi, r - two_pi, np.where(r < -math.pi, r + two_pi, r)).astype(np.float32)
        sp = np.sqrt(vel[:, 0] * vel[:, 0] + vel[:, 1] * vel[:, 1])
        gt[i, :, 0], gt[i, :, 1], gt[i, :, 2], gt[i, :, 3] = pos[:T1, 0], pos[:T1, 1], h[:T1], sp[:T1]
        gt_valid[i] = pos[:T1, 0] != np.float32(-10000.0)  # utils/sim.py:28 existence rule
        valid = np.asarray(o["valid"][:L], bool)
        if valid.any():
            last = int(np.nonzero(valid)[0][-1])
            target[i, 2], target[i, 3] = h[last], sp[last]
            dx, dy = pos[:, 0] - target[i, 0], pos[:, 1] - target[i, 1]
            dist = np.sqrt(dx * dx + dy * dy)
            moving[i] = bool(((sp > np.float32(speed_threshold)) | (dist > np.float32(moving_threshold)))[valid].any())
    segs, edge_polylines = [], []  # one pass over the roads: collision segments and road_edge_polylines() at once
    for road in scen["roads"]:
        g = road["geometry"]
        if road["type"] != "road_edge" or isinstance(g, dict):
            continue
        pts = _xy_array(g)
        edge_polylines.append(pts.astype(np.float64))
        if len(g) >= 2:
            segs.append(np.concatenate([pts[:-1], pts[1:]], axis=1))
    segs = np.concatenate(segs, axis=0) if segs else np.zeros((0, 4), np.float32)
    # goals (evaluators/evaluator.py:60-76), float64 views of float32 values
    goal = np.zeros((n, 4), np.float64)
    goal_idx = np.full(n, steps - 1, np.int64)  # step of the goal state (policy_evaluator.py:324-331)
    for i in range(n):
        gp = target[i, :2].astype(np.float64)
        gh, gs = float(target[i, 2]), float(target[i, 3])
        gone = np.where(gt_valid[i] == 0)[0]
        if len(gone) > 0:
            k = gone[0] - 1
            goal_idx[i] = k
            g64 = gt[i, k].astype(np.float64)
            if np.linalg.norm(g64[:2] - gp) > 0.0:
                gp, gh, gs = g64[:2], g64[2], g64[3]
        goal[i] = (gp[0], gp[1], gh, gs)
    goal_norm = np.linalg.norm(gt[:, 0, :2].astype(np.float64) - goal[:, :2], axis=1)
    return dict(n=n, gt=gt, gt_valid=gt_valid, size=size, moving=moving, goal=goal, goal_norm=goal_norm,
                segs=segs, goal_idx=goal_idx, edge_polylines=edge_polylines)


def interesting_pairs(parsed: dict, candidates, history_steps: int = 10, traj_len_threshold: int = 60,
                      goal_dist_threshold: float = 10.0, timestep_diff_threshold: int = 20):
    """The ordered vehicle pairs ``PolicyEvaluator.find_interesting_agent`` / ``find_interesting_pair`` draw from
    (evaluators/policy_evaluator.py:308-360,362-416): among the candidate (moving) vehicles, in scenario order, pairs
    (a, b) whose goals are distinct but less than ``goal_dist_threshold`` apart, reached within
    ``timestep_diff_threshold`` steps of each other, both with at least ``traj_len_threshold`` valid states after the
    history.  Row-major order of ``np.where`` over the candidate list, like the reference's ``valid_pairs``."""
    ids = [int(v) for v in candidates]
    if not ids:
        return []
    goals = parsed["goal"][ids, :2]
    goal_t = parsed["goal_idx"][ids] - history_steps
    long_enough = (parsed["gt_valid"][ids][:, history_steps:].astype(np.int64).sum(axis=1) >= traj_len_threshold).astype(np.int64)
    dists = np.linalg.norm(goals[None, :, :] - goals[:, None, :], 2, -1)
    mask = ((dists < goal_dist_threshold) * (dists > 0) * long_enough[:, None] * long_enough[None, :]
            * (np.abs(goal_t[:, None] - goal_t[None, :]) < timestep_diff_threshold))
    rows, cols = np.where(mask == 1)
    return [(ids[a], ids[b]) for a, b in zip(rows, cols)]


def road_arrays(preproc: dict):
    """road_points [P,100,3] float64 (x, y, exist) + one-hot road_types [P,8] -> xy, valid, type index."""
    rp = np.asarray(preproc["road_points"], np.float64)
    rt = np.asarray(preproc["road_types"])
    if rp.shape[0] == 0:
        return np.zeros((0, 100, 2)), np.zeros((0, 100), np.uint8), np.zeros(0, np.int8)
    types = np.where(rt.sum(-1) > 0, rt.argmax(-1), -1).astype(np.int8)
    return rp[:, :, :2].copy(), (rp[:, :, 2] != 0).astype(np.uint8), types


def road_edge_polylines(scen: dict):
    """The point lists the reference measures the signed distance to (``Evaluator.extract_road_edge_polylines``,
    evaluators/evaluator.py:143-158, on ``get_road_data``, utils/sim.py:67-73): every point of every ``road_edge`` road as
    the simulator holds it - ``RoadLine::geometry_points()`` are float32 (nocturne/cpp/include/road.h:113), so the file's
    doubles are rounded to float32 first.  Returns a list of [n_k, 2] float64 arrays."""
    out = []
    for road in scen["roads"]:
        g = road["geometry"]
        if isinstance(g, dict) or road["type"] != "road_edge":
            continue
        out.append(np.array([(q["x"], q["y"]) for q in g], np.float32).reshape(len(g), 2).astype(np.float64))
    return out


def initial_rtgs(cfg, preproc: dict):
    """[n, 3] un-normalised returns-to-go (position goal, vehicle-vehicle, vehicle-edge) of the LOGGED episode at t = 0:
    what a real_time_rewards policy without max_return / min_return starts from (``preproc_data['rtgs'][veh_idx, 0]``
    components 0, 3, 4; policy_evaluator.py:125-126).  The reference derives them when it loads the ``*_physics.pkl``:
    ``compute_rewards`` over the logged rewards and distances, then a reverse cumulative sum
    (datasets/rl_waymo/dataset_ctrl_sim.py:38-92 eval branch, dataset.py:239-275)."""
    w = cfg.dataset.waymo
    ag = np.asarray(preproc["ag_data"], np.float64)
    rew = np.asarray(preproc["ag_rewards"], np.float64)
    edge = np.asarray(preproc["veh_edge_dist_rewards"], np.float64)
    vv = np.asarray(preproc["veh_veh_dist_rewards"], np.float64)
    ex = ag[:, :, -1]
    goal = rew[:, :, 0] * w.pos_target_achieved_rew_multiplier
    if not w.remove_shaped_goal:
        goal = goal + (np.clip(rew[:, :, 3], w.pos_goal_shaped_min, w.pos_goal_shaped_max) - w.pos_goal_shaped_max) \
            * (1 / w.pos_goal_shaped_max)
    if w.remove_shaped_veh_reward:
        veh = -1 * rew[:, :, 6] * w.veh_veh_collision_rew_multiplier
    else:
        veh = vv - rew[:, :, 6] * w.veh_veh_collision_rew_multiplier
    if w.remove_shaped_edge_reward:
        road = -1 * rew[:, :, 7] * w.veh_edge_collision_rew_multiplier
    else:
        road = np.clip(np.abs(edge) * w.dist_to_road_edge_scaling_factor, 0, 5) / 5. - rew[:, :, 7] * w.veh_edge_collision_rew_multiplier
    allr = np.stack([goal * ex, veh * ex, road * ex], -1)
    return np.cumsum(allr[:, ::-1], axis=1)[:, ::-1][:, 0].copy()

