"""Synthetic Waymo-shaped scenes in the reference's on-disk formats.

There is no dataset in this environment, so every benchmark and parity input is synthesised.  A scene is emitted in
the two formats the reference evaluator reads:

  * the Nocturne scenario JSON consumed by ``Scenario::LoadScenario`` (nocturne/cpp/src/scenario.cc:207-264,
    893-1057): objects with 91-step position / heading (DEGREES) / velocity / valid tracks and a goalPosition,
    roads as typed polylines;
  * the ``*_physics.pkl`` dict read by ``RLWaymoDatasetCtRLSim.get_data`` (datasets/rl_waymo/dataset_ctrl_sim.py:
    38-97) of which only ``road_points`` [P,100,3] and ``road_types`` [P,8] influence a ctrl_sim policy
    (polylines chunked every 100 points exactly like ``get_roads``, datasets/rl_waymo/dataset.py:73-108).

Geometry: ``n_roads`` parallel straight roads, each 5 lanes of width 3.7 m (lane centre-lines,
type ``lane``), two ``road_edge`` polylines 3.7 m outside the outer lane centres and one ``road_line``; every
line is cut into ``n_chunks`` polylines of exactly 100 points that share their end points, so JSON road features
and model polylines are 1:1.  Vehicles sit on lane centres with >= 12 m headway, one common speed per lane
(+- 0.2 m/s) so that the constant-velocity ground truth never overlaps, heading {0, pi} + N(0, 0.003)
(0.02 rad would drift a 15 m/s vehicle 2.7 m sideways into the next lane over the 9 s episode).
"""
from __future__ import annotations

import json
import math
import os
import pickle

import numpy as np

LANE_W = 3.7
ROAD_TYPES = {"none": 0, "lane": 1, "road_line": 2, "road_edge": 3, "stop_sign": 4, "crosswalk": 5,
              "speed_bump": 6, "other": 7}  # utils/data.py:334-337


def make_scene(scene_id: int, n_vehicles: int = 64, n_roads: int = 4, n_chunks: int = 8, steps: int = 91,
               frac_short: float = 0.1, frac_parked: float = 0.0, road_spacing: float = 25.0,
               pts_spacing: float = 0.5, seed: int = 1234, world_offset: bool = True,
               speed_range=(3.0, 15.0), lane_ids=None):
    """Returns a dict with keys ``json`` (Nocturne schema) and ``preproc`` (pkl schema)."""
    rng = np.random.default_rng(seed + scene_id)
    ox, oy = (rng.uniform(-5000.0, 5000.0, size=2) if world_offset else (0.0, 0.0))
    n_pts = 100
    chunk_len = (n_pts - 1) * pts_spacing
    road_len = n_chunks * chunk_len
    x0 = -0.5 * road_len

    roads_json = []
    lane_specs = []  # (y, heading)
    for r in range(n_roads):
        yc = (r - 0.5 * (n_roads - 1)) * road_spacing
        lines = [("road_edge", yc - 2.5 * LANE_W - 0.5 * LANE_W), ("road_edge", yc + 2.5 * LANE_W + 0.5 * LANE_W),
                 ("road_line", yc + 0.5 * LANE_W)]
        for li in range(5):
            yl = yc + (li - 2) * LANE_W
            lines.append(("lane", yl))
            if lane_ids is None or li in lane_ids:  # lanes that may carry vehicles
                lane_specs.append((yl, 0.0 if li >= 3 else math.pi))
        # NB: lanes 0..2 (below the road_line) drive -x, lanes 3..4 drive +x
        for typ, y in lines:
            for c in range(n_chunks):
                xs = x0 + c * chunk_len + pts_spacing * np.arange(n_pts)
                roads_json.append({"type": typ,
                                   "geometry": [{"x": float(ox + x), "y": float(oy + y)} for x in xs]})

    # vehicles: fill lanes round-robin with >= 12 m headway
    per_lane = int(math.ceil(n_vehicles / len(lane_specs)))
    lane_speed = rng.uniform(speed_range[0], speed_range[1], size=len(lane_specs))
    lane_headway = 14.0 + rng.uniform(0.0, 6.0, size=len(lane_specs))
    objects = []
    order = rng.permutation(len(lane_specs) * per_lane)[:n_vehicles]
    dt = 0.1
    for slot in sorted(order.tolist()):
        lane, k = divmod(slot, per_lane)
        yl, hd = lane_specs[lane]
        headway = float(lane_headway[lane])
        span = headway * per_lane
        # start so that the 9 s of travel stays roughly centred on the road
        sgn = 1.0 if hd == 0.0 else -1.0
        xs0 = -sgn * (0.5 * span + 0.45 * lane_speed[lane] * 9.0) + sgn * (k * headway + rng.uniform(0.0, 2.0))
        heading = hd + rng.normal(0.0, 0.003)
        parked = rng.uniform() < frac_parked
        speed = 0.0 if parked else float(lane_speed[lane] + rng.uniform(-0.2, 0.2))
        length = float(rng.uniform(4.0, 5.5))
        width = float(rng.uniform(1.8, 2.2))
        last_valid = steps - 1
        if rng.uniform() < frac_short:
            last_valid = int(rng.integers(40, steps - 1))
        pos, head, vel, valid = [], [], [], []
        vx, vy = speed * math.cos(heading), speed * math.sin(heading)
        for t in range(steps):
            if t <= last_valid:
                pos.append({"x": float(ox + xs0 + vx * dt * t), "y": float(oy + yl + vy * dt * t)})
                head.append(float(math.degrees(heading)))
                vel.append({"x": float(vx), "y": float(vy)})
                valid.append(True)
            else:
                pos.append({"x": -10000.0, "y": -10000.0})
                head.append(-10000.0)
                vel.append({"x": -10000.0, "y": -10000.0})
                valid.append(False)
        objects.append({"type": "vehicle", "length": length, "width": width, "position": pos, "heading": head,
                        "velocity": vel, "valid": valid,
                        "goalPosition": dict(pos[last_valid])})
    name = f"synth_{scene_id}"
    scen = {"name": name + ".json", "objects": objects, "roads": roads_json, "tl_states": {}}
    return {"name": name, "json": scen, "preproc": preproc_from_json(scen)}


def make_replay_scene(i: int):
    """Scene i of the BASELINE config-4 workload ("Waymo val_interactive-shaped replay"): varied vehicle counts (4..64),
    road layouts, parked and short-lived vehicles; in every fifth scene the first vehicle leaves its lane at 20 degrees,
    runs into its neighbours and crosses a road edge, so that collisions, contact response and off-road events occur."""
    sc = make_scene(5000 + i, n_vehicles=4 + (7 * i) % 61, n_roads=2 + i % 3, n_chunks=3 + i % 5, frac_short=0.25,
                    frac_parked=0.2 if i % 2 else 0.0)
    if i % 5 == 0:
        o = sc["json"]["objects"][0]
        th = math.radians(o["heading"][0]) + 0.35
        sp = max(6.0, math.hypot(o["velocity"][0]["x"], o["velocity"][0]["y"]))
        x0, y0 = o["position"][0]["x"], o["position"][0]["y"]
        for t, ok in enumerate(o["valid"]):
            if ok:
                o["position"][t] = {"x": x0 + sp * 0.1 * t * math.cos(th), "y": y0 + sp * 0.1 * t * math.sin(th)}
                o["heading"][t] = math.degrees(th)
                o["velocity"][t] = {"x": sp * math.cos(th), "y": sp * math.sin(th)}
        last = max(t for t, ok in enumerate(o["valid"]) if ok)
        o["goalPosition"] = dict(o["position"][last])
    return sc


def replay_scene_summary(pos, heading, existence, reward, gt_pos):
    """Per-scene figures of a log-replay episode (BASELINE config 4): vehicle-steps in a vehicle-vehicle collision,
    off-road vehicle-steps, vehicles that ever collided / left the road, ADE against the logged positions, and an
    order-independent checksum of the simulated positions.  Arrays [n, 91, ...] in the trace layout."""
    ex = existence.astype(bool)
    cv, ce = (reward[:, :, 6] > 0) & ex, (reward[:, :, 7] > 0) & ex
    err = np.linalg.norm(pos.astype(np.float64) - gt_pos.astype(np.float64), axis=-1)
    return {"n": int(pos.shape[0]), "veh_steps": int(ex.sum()), "coll_steps": int(cv.sum()), "off_steps": int(ce.sum()),
            "coll_veh": int(cv.any(1).sum()), "off_veh": int(ce.any(1).sum()),
            "ade": float(err[ex].mean()) if ex.any() else 0.0,
            "pos_sum": float(np.abs(pos.astype(np.float64))[ex].sum()), "head_sum": float(np.abs(heading.astype(np.float64))[ex].sum())}


def preproc_from_json(scen, n_pts: int = 100):
    """road_points [P,100,3] (x, y, exist) and road_types [P,8], chunked like datasets/rl_waymo/dataset.py:73-108."""
    pts, types = [], []
    for road in scen["roads"]:
        geom = road["geometry"]
        onehot = np.eye(8)[ROAD_TYPES.get(road["type"], 7)]
        if isinstance(geom, dict):
            pts.append(np.tile(np.array([geom["x"], geom["y"], 1.0]), (n_pts, 1)))
            types.append(onehot)
            continue
        for s in range(0, len(geom), n_pts):
            chunk = geom[s:s + n_pts]
            arr = np.zeros((n_pts, 3))
            arr[:len(chunk), 0] = [p["x"] for p in chunk]
            arr[:len(chunk), 1] = [p["y"] for p in chunk]
            arr[:len(chunk), 2] = 1.0
            pts.append(arr)
            types.append(onehot)
    n_obj = len(scen["objects"])
    return {
        "idx": 0, "num_agents": n_obj,
        "road_points": np.array(pts), "road_types": np.array(types),
        # the remaining keys only feed `rtgs`, which a ctrl_sim policy ignores unless real_time_rewards is set
        "ag_data": np.zeros((n_obj, 90, 8)), "ag_actions": np.zeros((n_obj, 90, 2)), "ag_types": np.zeros((n_obj, 5)),
        "last_exist_timesteps": np.zeros(n_obj), "ag_rewards": np.zeros((n_obj, 90, 8)),
        "veh_edge_dist_rewards": np.zeros((n_obj, 90)), "veh_veh_dist_rewards": np.zeros((n_obj, 90)),
        "filtered_ag_ids": list(range(n_obj)), "ag_goals": np.zeros((n_obj, 90, 5)),
    }


def write_dataset(root: str, scenes) -> dict:
    """Lay scenes out the way the reference evaluator expects them under ``root`` and return the path settings.

    <root>/test_filenames.pkl, <root>/json/<name>.json, <root>/preprocess/test/<name>_physics.pkl
    (evaluators/policy_evaluator.py:33-35, evaluators/evaluator.py:44-57).
    """
    jdir = os.path.join(root, "json")
    pdir = os.path.join(root, "preprocess", "test")
    os.makedirs(jdir, exist_ok=True)
    os.makedirs(pdir, exist_ok=True)
    names = []
    for sc in scenes:
        with open(os.path.join(jdir, sc["name"] + ".json"), "w") as f:
            json.dump(sc["json"], f)
        with open(os.path.join(pdir, sc["name"] + "_physics.pkl"), "wb") as f:
            pickle.dump(sc["preproc"], f)
        names.append(sc["name"] + ".json")
    with open(os.path.join(root, "test_filenames.pkl"), "wb") as f:
        pickle.dump({"test_filenames": names}, f)
    return {"dataset_root": root, "nocturne_waymo_val_folder": jdir,
            "preprocess_dir": os.path.join(root, "preprocess")}
