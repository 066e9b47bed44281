"""Planner-vs-adversary evaluation throughput on BASELINE config-2 shaped scenes: in every scene the
planner drives vehicle 0 and the adversary vehicle 1, the other 62 vehicles are log-replayed.

    python tools/planner_adversary_bench.py [--scenes 256] [--cat]

Prints one JSON line: controlled-agent-steps/s (2 per scene and step; 1 with --cat), ms per simulated step, metrics."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ctrlsim_b200.config import default_config
from ctrlsim_b200.evaluator import B200Policy
from ctrlsim_b200.model import DeviceModel
from ctrlsim_b200.planner_adversary import B200PlannerAdversaryEvaluator, CatAdversary
from ctrlsim_b200.synth import make_scene
from ctrlsim_b200.weights import make_weights

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=256)
ap.add_argument("--cat", action="store_true")
args = ap.parse_args()
cfg = default_config()
dev = torch.device("cuda:0")
weights = make_weights(cfg, seed=0)
scenes = [make_scene(i) for i in range(args.scenes)]
tilt = lambda g, v, e: {"tilt": True, "goal_tilt": g, "veh_veh_tilt": v, "veh_edge_tilt": e}
planner = B200Policy(cfg, "synthetic", DeviceModel(cfg, weights, dev), tilt_dict=tilt(10, 10, 10), seed=1, chunk_groups=256)
if args.cat:
    adversary = CatAdversary()
    trajs = [np.array([[p["x"], p["y"]] for p in sc["json"]["objects"][1]["position"]])[:91] for sc in scenes]
else:
    adversary = B200Policy(cfg, "synthetic", DeviceModel(cfg, weights, dev), tilt_dict=tilt(0, -10, 0), seed=2,
                           chunk_groups=256)
    trajs = None
ev = B200PlannerAdversaryEvaluator(cfg, planner, adversary, scenes=scenes, pairs=[(0, 1)] * len(scenes), adv_trajs=trajs)
ev.build_batch()
ev.rollout(max_steps=3)  # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
metrics, _ = ev.evaluate_planner_adversary()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
controlled = (1 if args.cat else 2) * len(scenes)
print(json.dumps({"metric": "controlled-agent-steps/s (planner vs adversary)", "value": controlled * 90 / dt,
                  "scenes": len(scenes), "adversary": "cat" if args.cat else "ctrl_sim", "ms_per_step": 1000 * dt / 90,
                  "seconds": dt, "metrics": metrics}))
