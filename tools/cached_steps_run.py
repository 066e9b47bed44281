"""Steps 0..T-1 of one episode (the cached phase: prefix + map caches, incremental decode) for an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python tools/cached_steps_run.py 32 14
   python tools/ncu_shares.py X.csv        (shares of the last complete step)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlsim_b200.config import default_config
from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
from ctrlsim_b200.model import DeviceModel
from ctrlsim_b200.weights import make_weights

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 14
cfg = default_config()
import bench
scenes, ids = bench.make_scenes(n_scenes, 0, 1)
pol = B200Policy(cfg, "synthetic", DeviceModel(cfg, make_weights(cfg, seed=0), "cuda:0"), seed=0, chunk_groups=256)
ev = B200PolicyEvaluator(cfg, pol, scenes=scenes, scene_ids=ids)
b = ev.build_batch(eval_threshold=64)
ev.rollout(b, max_steps=steps)
torch.cuda.synchronize()
print("groups last step", pol.groups_last_step)
