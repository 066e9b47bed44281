"""Diagnostic: planner-vs-adversary rollouts of the committed fixtures on the GPU, listing every sampled bin that differs
from the unmodified reference (tests/golden/planner_adversary_*.npz) and the final metrics side by side."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import load_planner_adversary_golden
from ctrlsim_b200.config import default_config
from ctrlsim_b200.evaluator import B200Policy
from ctrlsim_b200.model import DeviceModel
from ctrlsim_b200.planner_adversary import B200PlannerAdversaryEvaluator, CatAdversary
from ctrlsim_b200.synth import make_scene
from ctrlsim_b200.weights import make_weights
cfg = default_config(); dev = torch.device("cuda:0")
for name in ("policies", "cat"):
    recs, spec, ref = load_planner_adversary_golden(name)
    scenes = [make_scene(**s) for s in spec["scenes"]]
    weights = make_weights(cfg, **spec["weights"])
    tilt = lambda t: {"tilt": True, "goal_tilt": t[0], "veh_veh_tilt": t[1], "veh_edge_tilt": t[2]}
    planner = B200Policy(cfg, "synthetic", DeviceModel(cfg, weights, dev), tilt_dict=tilt(spec["tilts_planner"]), seed=spec["seeds"][0])
    adversary = CatAdversary() if spec["cat"] else B200Policy(cfg, "synthetic", DeviceModel(cfg, weights, dev), tilt_dict=tilt(spec["tilts_adversary"]), seed=spec["seeds"][1])
    ev = B200PlannerAdversaryEvaluator(cfg, planner, adversary, scenes=scenes, pairs=[tuple(p) for p in spec["pairs"]], adv_trajs=[r["adv_pos"] for r in recs])
    m, _ = ev.evaluate_planner_adversary()
    tr = ev.batch.trace()
    for s, g in enumerate(recs):
        n = g["pos"].shape[0]
        for role, view in (("planner", ev.view_planner), ("adversary", ev.view_adversary)):
            if view is None: continue
            a = view.t["tr_rtg_idx"][s, :n, :90].cpu().numpy().transpose(1, 0, 2); b = g[f"{role}_rtg_idx"]
            bad = np.argwhere(a != b)
            print(name, s, role, "rtg mismatches", len(bad), bad[:8].tolist())
            for t, v, c in bad[:8]:
                print("   t v c", t, v, c, "gpu", a[t, v], "gold", b[t, v], "exist", g["existence"][v, t], "members", g[f"{role}_members"][t][:5])
            aa = view.t["tr_act_idx"][s, :n, :90].cpu().numpy().T
            print(name, s, role, "act mismatches", np.argwhere(aa != g[f"{role}_act_idx"])[:5].tolist())
        ex = g["existence"][:, :90].astype(bool)
        print(name, s, "dpos", np.abs(tr["tr_pos"][s, :n, :90].astype(np.float64) - g["pos"][:, :90])[ex].max(),
              "dacc", np.abs(tr["tr_action"][s, :n, :90] - np.stack([g["accel"], g["steer"]], -1)[:, :90])[ex].max())
    print(name, {k: (round(m[k], 6), round(v, 6)) for k, v in ref.items()})
