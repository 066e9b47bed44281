"""BASELINE config 3: controllability sweep over the three reward-tilt exponents, scenes sharded over the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/tilt_sweep.py --scenes 2048
    python tools/tilt_sweep.py --scenes 64            # single GPU

One evaluation (90-step rollout of every scene + ONE all-reduce of the 1608-double summary) per sweep point; the grid
(each exponent in {-25, -10, 0, +10, +25} one at a time, 13 points) is SURVEY 8(d)'s choice - the reference documents
the mechanism and the values 0 / +10 / -10 (cfgs/policy/ctrl_sim.yaml:6-8, ctrl_sim_adversary.yaml:7).  Rank 0 prints
one JSON line per point with the reference's metric keys and the agent-steps/s of that evaluation."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from ctrlsim_b200.config import default_config
from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
from ctrlsim_b200.model import DeviceModel
from ctrlsim_b200.synth import make_scene
from ctrlsim_b200.weights import make_weights

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=64, help="scenes of the whole job (sharded over ranks)")
ap.add_argument("--values", type=float, nargs="*", default=[-25.0, -10.0, 10.0, 25.0])
ap.add_argument("--chunk", type=int, default=256)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = default_config()
model = DeviceModel(cfg, make_weights(cfg, seed=0), dev)
# every rank generates only the scenes it owns (scene i -> rank i % world, the evaluator's own sharding rule; all 64
# vehicles of every scene are evaluated, so no evaluated-set draw depends on the other ranks' scenes)
mine = list(range(rank, args.scenes, world))
scenes = [make_scene(i) for i in mine]
points = [("none", (0.0, 0.0, 0.0))] + [(f"{n}={v:+g}", tuple(v if k == j else 0.0 for k in range(3)))
                                        for j, n in enumerate(("goal", "veh_veh", "veh_edge")) for v in args.values]
batch = None
for name, (tg, tv, te) in points:
    pol = B200Policy(cfg, "synthetic", model, tilt_dict={"tilt": True, "goal_tilt": tg, "veh_veh_tilt": tv, "veh_edge_tilt": te},
                     seed=0, chunk_groups=args.chunk)
    ev = B200PolicyEvaluator(cfg, pol, scenes=scenes, scene_ids=mine)
    if batch is None:  # the scene batch is uploaded once; every sweep point resets its dynamic state (Policy.reset)
        ev.rank, ev.world = 0, 1
        batch = ev.build_batch(eval_threshold=64)
    ev.rank, ev.world, ev.batch = rank, world, batch
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    metrics, _ = ev.evaluate_policy()          # rollout + metrics kernel + the one all-reduce
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    w = torch.tensor([wall], dtype=torch.float64, device=dev)
    n = torch.tensor([float(batch.n_evaluated())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(n)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"point": name, "tilts": [tg, tv, te], "n_gpus": world, "scenes": args.scenes,
                          "agent_steps_per_s": n.item() * cfg.nocturne.steps / w.item(), "wall_s": w.item(), "metrics": metrics}), flush=True)
    del ev, pol
if world > 1:
    dist.destroy_process_group()
