"""Is a simulator step host-bound?  Wall time the HOST spends inside update_state / predict / act (enqueue only, no
synchronize) against the device time of the step (CUDA events), per phase."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ctrlsim_b200.config import default_config
from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
from ctrlsim_b200.model import DeviceModel
from ctrlsim_b200.weights import make_weights

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = default_config()
scenes, ids = bench.make_scenes(n_scenes, 0, 1)
pol = B200Policy(cfg, "synthetic", DeviceModel(cfg, make_weights(cfg, seed=0), "cuda:0"), seed=0, chunk_groups=256)
ev = B200PolicyEvaluator(cfg, pol, scenes=scenes, scene_ids=ids)
b = ev.build_batch(eval_threshold=64)
pol.reset(b)
rows = []
for t in range(40):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    e0.record()
    pol.update_state(b, t)
    h1 = time.perf_counter()
    pol.predict(b, t)
    h2 = time.perf_counter()
    pol.act(b, t)
    e1.record()
    h3 = time.perf_counter()
    torch.cuda.synchronize()
    rows.append((t, e0.elapsed_time(e1), (h1 - h0) * 1e3, (h2 - h1) * 1e3, (h3 - h2) * 1e3))
for t, dev_ms, a, p, c in rows:
    if t in (1, 5, 12, 20, 31, 32, 35, 39):
        print(f"t={t:2d} device {dev_ms:8.2f} ms | host: update_state {a:6.2f}  predict {p:7.2f}  act {c:5.2f} ms  (groups {pol.groups_last_step})")
