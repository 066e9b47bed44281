"""ORACLE (test infrastructure only): fixtures tests/golden/planner_adversary_*.npz from the REFERENCE itself.

    python -m oracle.make_golden_planner_adversary [--only NAME]

Runs the unmodified ``PlannerAdversaryEvaluator`` (oracle/ref_harness.run_reference_planner_adversary) on synthetic
scenes - build container only (needs /root/reference and oracle/_ref).  Scenes are of the "sparse" kind of
oracle/make_golden.py (one vehicle per road, near-deterministic coasting through a large still_bias) so that no Box2D
contact occurs and the whole 90-step episode, including the end-of-episode metrics, is comparable.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")

SPARSE = dict(n_vehicles=6, n_roads=6, n_chunks=4, lane_ids=[3], frac_short=0.34, speed_range=(5.0, 12.0))
FIXTURES = {
    # planner (tilts +10) and adversary (veh_veh_tilt -10) are two CtRL-Sim policies with their own sampler seeds
    "policies": dict(scenes=[dict(scene_id=21, **SPARSE), dict(scene_id=22, **SPARSE)], pairs=[(0, 1), (2, 4)],
                     weights=dict(seed=2, still_bias=12.0), seeds=(1, 2), tilts_planner=(10, 10, 10),
                     tilts_adversary=(0, -10, 0), cat=False),
    # scripted CAT adversary: the logged track of the adversary, stretched 6 % ahead and pushed 1.2 m sideways
    "cat": dict(scenes=[dict(scene_id=23, **SPARSE)], pairs=[(1, 3)], weights=dict(seed=2, still_bias=12.0),
                seeds=(3, 0), tilts_planner=(10, 10, 10), tilts_adversary=(0, 0, 0), cat=True),
}


def scripted_positions(scene, obj_idx, cat):
    pos = np.array([[p["x"], p["y"]] for p in scene["json"]["objects"][obj_idx]["position"]], np.float64)[:91]
    if not cat:
        return pos
    k = np.linspace(0.0, 1.0, len(pos))[:, None]
    return pos[0] + (pos - pos[0]) * 1.06 + np.array([0.0, 1.2]) * k


def pack(rec):
    out = {k: v for k, v in rec.items() if isinstance(v, np.ndarray)}
    for role in ("planner", "adversary"):
        sub = rec[role]
        if sub is None:
            continue
        out[f"{role}_act_idx"], out[f"{role}_rtg_idx"], out[f"{role}_rtgs"] = sub["act_idx"], sub["rtg_idx"], sub["rtgs"]
        steps = len(sub["groups"])
        focal = -np.ones(steps, np.int32)
        members = -np.ones((steps, 24), np.int32)
        for t, gs in enumerate(sub["groups"]):
            assert len(gs) <= 1
            for d in gs:
                focal[t], members[t] = d["focal"], d["members"]
        out[f"{role}_focal"], out[f"{role}_members"] = focal, members
    out["ego_adv"] = np.array([rec["ego"], rec["adv"]], np.int32)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    sys.path.insert(0, ROOT)
    from ctrlsim_b200.config import default_config
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    from oracle.ref_harness import run_reference_planner_adversary

    for name, spec in FIXTURES.items():
        if args.only and name != args.only:
            continue
        t0 = time.time()
        scenes = [make_scene(**s) for s in spec["scenes"]]
        adv_pos = [scripted_positions(sc, p[1], spec["cat"]) for sc, p in zip(scenes, spec["pairs"])]
        weights = make_weights(default_config(), **spec["weights"])
        metrics, recs = run_reference_planner_adversary(
            scenes, spec["pairs"], adv_pos, weights=weights, seeds=spec["seeds"], tilts_planner=spec["tilts_planner"],
            tilts_adversary=spec["tilts_adversary"], cat=spec["cat"])
        out = {}
        for k, rec in enumerate(recs):
            contact = np.where((rec["reward"][:, :, 6] * rec["existence"]).any(axis=0))[0]
            print(f"[golden] {name} scene {k}: first vehicle-vehicle contact at step {contact[:1]}", flush=True)
            for key, v in pack(rec).items():
                out[f"s{k}_{key}"] = v
            out[f"s{k}_adv_pos"] = adv_pos[k]
        out["metrics_json"] = np.frombuffer(json.dumps(metrics).encode(), dtype=np.uint8)
        out["spec_json"] = np.frombuffer(json.dumps(spec).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(GOLDEN, f"planner_adversary_{name}.npz"), **out)
        print(f"[golden] {name}: {time.time() - t0:.1f}s metrics={metrics}", flush=True)


if __name__ == "__main__":
    main()
