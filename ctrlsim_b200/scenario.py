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


def parse_scenario(scen: dict, steps: int = 90, moving_threshold: float = 0.2, speed_threshold: float = 0.05):
    objs = [o for o in scen["objects"] if bool(o["valid"][0]) and o["type"] == "vehicle"]
    n, T1 = len(objs), steps + 1
    gt = np.zeros((n, T1, 4), np.float32)
    gt_valid = np.zeros((n, T1), np.uint8)
    size = np.zeros((n, 2), np.float32)
    target = np.zeros((n, 4), np.float32)
    moving = np.zeros(n, bool)
    for i, o in enumerate(objs):
        if len(o["position"]) < T1:
            raise ValueError(f"object {i}: {len(o['position'])} states, the evaluator needs {T1}")
        size[i] = (o["length"], o["width"])
        gp = o.get("goalPosition", {"x": 0.0, "y": 0.0})
        target[i, :2] = (gp["x"], gp["y"])
        for t in range(len(o["position"])):
            x, y = np.float32(o["position"][t]["x"]), np.float32(o["position"][t]["y"])
            h = _normalize_angle_f32(o["heading"][t])
            vx, vy = np.float32(o["velocity"][t]["x"]), np.float32(o["velocity"][t]["y"])
            sp = np.sqrt(np.float32(vx * vx + vy * vy))
            if t < T1:
                gt[i, t] = (x, y, h, sp)
                gt_valid[i, t] = x != np.float32(-10000.0)  # utils/sim.py:28 existence rule
            if bool(o["valid"][t]):
                target[i, 2], target[i, 3] = h, sp
                dx, dy = x - target[i, 0], y - target[i, 1]
                dist = np.sqrt(np.float32(dx * dx + dy * dy))
                if sp > np.float32(speed_threshold) or dist > np.float32(moving_threshold):
                    moving[i] = True
    segs = []
    for road in scen["roads"]:
        g = road["geometry"]
        if road["type"] != "road_edge" or isinstance(g, dict):
            continue
        for k in range(len(g) - 1):
            segs.append((g[k]["x"], g[k]["y"], g[k + 1]["x"], g[k + 1]["y"]))
    # goals (evaluators/evaluator.py:60-76), float64 views of float32 values
    goal = np.zeros((n, 4), np.float64)
    for i in range(n):
        gp = target[i, :2].astype(np.float64)
        gh, gs = float(target[i, 2]), float(target[i, 3])
        gone = np.where(gt_valid[i] == 0)[0]
        if len(gone) > 0:
            k = gone[0] - 1
            g64 = gt[i, k].astype(np.float64)
            if np.linalg.norm(g64[:2] - gp) > 0.0:
                gp, gh, gs = g64[:2], g64[2], g64[3]
        goal[i] = (gp[0], gp[1], gh, gs)
    goal_norm = np.linalg.norm(gt[:, 0, :2].astype(np.float64) - goal[:, :2], axis=1)
    return dict(n=n, gt=gt, gt_valid=gt_valid, size=size, moving=moving, goal=goal, goal_norm=goal_norm,
                segs=np.asarray(segs, np.float32).reshape(-1, 4))


def road_arrays(preproc: dict):
    """road_points [P,100,3] float64 (x, y, exist) + one-hot road_types [P,8] -> xy, valid, type index."""
    rp = np.asarray(preproc["road_points"], np.float64)
    rt = np.asarray(preproc["road_types"])
    if rp.shape[0] == 0:
        return np.zeros((0, 100, 2)), np.zeros((0, 100), np.uint8), np.zeros(0, np.int8)
    types = np.where(rt.sum(-1) > 0, rt.argmax(-1), -1).astype(np.int8)
    return rp[:, :, :2].copy(), (rp[:, :, 2] != 0).astype(np.uint8), types
