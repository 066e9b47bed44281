"""Small rollout for compute-sanitizer (tools/sanitize.sh): one scene, 35 steps, so that the cached steps (t < 32), the
first full-window steps (t >= 32), log replay, the simulator and the metrics kernels all launch at least once."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlsim_b200.config import default_config
from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
from ctrlsim_b200.model import DeviceModel
from ctrlsim_b200.synth import make_scene
from ctrlsim_b200.weights import make_weights

cfg = default_config()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 35
pol = B200Policy(cfg, "synthetic", DeviceModel(cfg, make_weights(cfg, seed=0), "cuda:0"), seed=0)
ev = B200PolicyEvaluator(cfg, pol, scenes=[make_scene(11, n_vehicles=6, n_roads=1, n_chunks=4),
                                           make_scene(12, n_vehicles=30, n_roads=2, n_chunks=3)])
b = ev.build_batch(eval_threshold=64)
ev.rollout(b, max_steps=steps)
s = ev.summarize(b)
torch.cuda.synchronize()
tr = b.trace()
coll = int(((tr["tr_reward"][..., 6] > 0) & (tr["tr_exist"] > 0)).sum())
print(f"sanitize target ok: {b.n_evaluated()} evaluated vehicles, {steps} steps, groups last step {pol.groups_last_step}, "
      f"{coll} vehicle-steps in a vehicle-vehicle collision (Box2D contact response exercised), contact overflow {b.contact_overflow()}")
