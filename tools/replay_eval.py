"""BASELINE config 4: 1000 Waymo val_interactive-SHAPED scenes, every vehicle log-replayed (inverse bicycle model ->
FreeCar -> Box2D world step with contact response -> collision / off-road flags -> rewards) as ONE GPU batch; per-scene
collision / off-road / ADE figures, compared with the CPU reference table of tests/golden/replay_oracle.json
(oracle/make_replay_table.py: the first scenes of the same workload through the C restatement of the reference
simulator, which is bit-identical to the real nocturne_cpp).

    python tools/replay_eval.py [--scenes 1000] > profiles/r02_replay_1000.json
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ctrlsim_b200.config import default_config
from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
from ctrlsim_b200.model import DeviceModel
from ctrlsim_b200.synth import make_replay_scene, replay_scene_summary
from ctrlsim_b200.weights import make_weights

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=1000)
args = ap.parse_args()
dev = torch.device("cuda:0")
cfg = default_config()
t0 = time.perf_counter()
scenes = [make_replay_scene(i) for i in range(args.scenes)]
t_gen = time.perf_counter() - t0
pol = B200Policy(cfg, "synthetic", DeviceModel(cfg, make_weights(cfg, seed=0), dev), seed=0)
ev = B200PolicyEvaluator(cfg, pol, scenes=scenes)
t0 = time.perf_counter()
b = ev.build_batch(eval_threshold=0, keep_replay_only=True)   # nobody is policy-controlled
torch.cuda.synchronize()
t_build = time.perf_counter() - t0
ev.rollout(b)                                                  # warm-up episode
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ev.rollout(b)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
tr = b.trace()
gt = b.t["gt"].cpu().numpy()
rows = []
for s in range(args.scenes):
    n = int(tr["n_veh"][s])
    rows.append(replay_scene_summary(tr["tr_pos"][s, :n], tr["tr_heading"][s, :n], tr["tr_exist"][s, :n], tr["tr_reward"][s, :n],
                                     gt[s, :n, :, :2]))
tot = {k: sum(r[k] for r in rows) for k in ("n", "veh_steps", "coll_steps", "off_steps", "coll_veh", "off_veh")}
out = {"workload": f"{args.scenes} synthetic Waymo val_interactive-shaped scenes (ctrlsim_b200.synth.make_replay_scene), every vehicle "
                   "log-replayed for 90 steps, one GPU batch", "gpu_ms_per_episode": ms,
       "vehicle_steps_per_s": tot["veh_steps"] / (ms * 1e-3), "host_scene_generation_s": t_gen, "host_parse_and_upload_s": t_build,
       "totals": tot, "collision_rate_vehicles": tot["coll_veh"] / tot["n"], "offroad_rate_vehicles": tot["off_veh"] / tot["n"],
       "ade_m_mean_over_scenes": float(np.mean([r["ade"] for r in rows])), "contact_overflow": b.contact_overflow()}
ref_path = os.path.join(ROOT, "tests", "golden", "replay_oracle.json")
if os.path.exists(ref_path):
    ref = json.load(open(ref_path))["rows"][:args.scenes]
    cmp_ = {"scenes_compared": len(ref), "identical_collision_steps": 0, "identical_offroad_steps": 0, "identical_position_checksum": 0,
            "max_ade_diff_m": 0.0, "scenes_with_contacts": 0}
    for r, g in zip(ref, rows):
        cmp_["identical_collision_steps"] += r["coll_steps"] == g["coll_steps"] and r["coll_veh"] == g["coll_veh"]
        cmp_["identical_offroad_steps"] += r["off_steps"] == g["off_steps"] and r["off_veh"] == g["off_veh"]
        cmp_["identical_position_checksum"] += r["pos_sum"] == g["pos_sum"] and r["head_sum"] == g["head_sum"]
        cmp_["max_ade_diff_m"] = max(cmp_["max_ade_diff_m"], abs(r["ade"] - g["ade"]))
        cmp_["scenes_with_contacts"] += r["coll_steps"] > 0
    cmp_["oracle_totals"] = {k: sum(r[k] for r in ref) for k in ("veh_steps", "coll_steps", "off_steps", "coll_veh", "off_veh")}
    cmp_["gpu_totals_same_scenes"] = {k: sum(r[k] for r in rows[:len(ref)]) for k in ("veh_steps", "coll_steps", "off_steps", "coll_veh", "off_veh")}
    out["vs_cpu_reference"] = cmp_
print(json.dumps(out, indent=1))
