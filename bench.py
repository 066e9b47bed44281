"""Benchmark of the closed-loop rollout hot path (BASELINE.json metric: agent-steps/s; encoder-attn HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--scenes S] [--impl b200|reference]

Workload (BASELINE.json configs[1], per GPU): S=256 synthetic Waymo-shaped scenes x 64 policy-controlled vehicles x
256 polylines, RTG-conditioned autoregressive policy, 90-step episodes (evaluators/policy_evaluator.py:514-557).
A "step" is one simulator step of the whole scene batch: observe -> focal grouping -> tokenise -> two-pass network ->
RTG/action sampling -> log-replay/act -> physics + collision.

An episode has two very different phases: steps 0..31 (the 32-step window still starts at t=0, prefix + map caches
apply, ~6 % of the episode's time) and steps 32..89 (window slides, everything is recomputed).  Whatever K is, the
timed steps are drawn from both phases IN EPISODE PROPORTION (58 : 32): K = 90 q + r is timed as q whole episodes plus
round(r 58/90) full-window steps (t >= 32) plus the remaining steps of the cached phase, ending at t = 31.  Each timed
segment is bracketed by barrier + synchronize and timed with CUDA events; between segments the episode is
fast-forwarded untimed.  W warm-up steps run untimed in front of each phase's segment.  So `value` is the whole-episode
throughput for every K (the `phases` object gives the per-step time of either phase).  For r > 0 it is slightly
CONSERVATIVE: a cached step attends to t x 72 cached keys, so its cost grows linearly with t (13.6 ms at t = 1, 26.6 ms
at t = 31 for 64 scenes, `tools/host_time_probe.py`), and the cached segment ends at t = 31 - the most expensive cached
steps stand for the phase (whole-episode truth at 256 scenes: ~3 % above the K = 20 figure; K = 90 has no such bias).  Scenes shard over ranks with
no data-path collective ("weak": per-GPU work fixed); the only collective is the summary all-reduce at the end of an
evaluation, exercised in the e2e leg.

  value          agent-steps/s, state resident in HBM, CUDA-event timed, max over ranks
  e2e            the same metric through the public API from HOST scene arrays: B200PolicyEvaluator.evaluate_policy()
                 = pinned host -> device upload of the scene batch + reset + 90 steps + metrics kernel + all-reduce +
                 device -> host read of the summary and of the full per-vehicle trace
  roofline       dominant kernel class of the step, timed live with CUDA events on the launching stream
  encoder_attn   the polyline-pooling attention kernel (HBM-bound), the kernel BASELINE.json's metric names
  cpu_baseline   the reference's algorithm on this box's host cores (oracle port: two full B=1 2304-token forwards per
                 focal group per step + C simulator), process-parallel: one single-threaded process per core, one scene
                 of the same workload each, a stated prefix of the episode
  gpu_torch_baseline  the same port's network on cuda:0 with stock torch ops at B=1 (the reference's GPU cost structure,
                 policies/autoregressive_policy.py:183-210): forward time only, so an upper bound for that path
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T_WIN = 32  # context length: steps t < 32 are the cached phase
NPTS = 100  # points per polyline


def workload_name(scenes, wide):
    geom = "wide model caps A=64/P=256 (focal groups of up to 64 agents; the 60 m context radius still splits a scene into ~10 groups)" if wide else "reference-default caps A=24/P=200 (~11.6 focal groups per scene)"
    return (f"{scenes} synthetic 64-agent/256-polyline scenes per GPU, 90-step episodes, RTG-conditioned autoregressive "
            f"policy, {geom}")


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "hbm_src": "measured", "tf": d["bf16_tflops_sustained"], "tf_src": "measured (sustained bf16 cuBLAS)"}
    return {"hbm": 6650.0, "hbm_src": "fallback", "tf": 1400.0, "tf_src": "fallback"}


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep files)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,power.limit"

    def __init__(self, index=0):
        self.rows, self.p, self.index = [], None, index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                       str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        return len(self.rows)

    def stop(self, spans=None):
        """spans: list of (row_begin, row_end) index pairs of the timed segments; only those samples count."""
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        rows = self.rows if not spans else [r for a, b in spans for r in self.rows[a:max(b, a + 1)]]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        def fl(r, i):
            try:
                return float(r[i])
            except (IndexError, ValueError):
                return None
        pw = sorted(v for v in (fl(r, 6) for r in rows) if v is not None)
        lim = [v for v in (fl(r, 7) for r in rows) if v is not None]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "power_w": pw[len(pw) // 2] if pw else None, "power_limit_w": max(lim) if lim else None}


def make_scenes(n, first_id, stride, **kw):
    from ctrlsim_b200.synth import make_scene
    return [make_scene(first_id + i * stride, **kw) for i in range(n)], [first_id + i * stride for i in range(n)]


# ---- the reference's CPU cost structure on host cores --------------------------------------------------------------
def cpu_worker(scene_id, steps, wide):
    """One single-threaded process = one scene of the workload, first `steps` steps of its episode."""
    import torch
    torch.set_num_threads(1)
    from ctrlsim_b200.config import default_config
    from ctrlsim_b200.weights import make_weights
    from oracle.model_port import ModelPort
    from oracle.policy_port import RolloutPort
    cfg = default_config(wide=wide)
    weights = make_weights(cfg, seed=0)
    scenes, ids = make_scenes(1, scene_id, 1)
    port = RolloutPort(cfg, ModelPort(cfg, weights), seed=0, eval_threshold=64)
    print("READY", flush=True)
    sys.stdin.readline()  # all workers start together
    t0 = time.perf_counter()
    rec = port.run_scene(ids[0], scenes[0]["json"], scenes[0]["preproc"], max_steps=steps)
    wall = time.perf_counter() - t0
    print(json.dumps({"agents": len(rec["evaluated"]), "steps": steps, "wall": wall, "forwards": port.n_forwards}), flush=True)


def cpu_reference_run(steps, wide=False, n_proc=None):
    """N_proc = host cores, one single-threaded process per core, each running whole scenes of
    the workload (here: the first `steps` steps of one scene each).  Returns (agent-steps/s, cores, sample, detail)."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    n_proc = n_proc or max(1, cores)
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker", str(10_000 + i), "--cpu-steps",
                               str(steps)] + (["--wide"] if wide else []), stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                              text=True, env=env, cwd=ROOT) for i in range(n_proc)]
    for p in procs:
        line = p.stdout.readline()
        assert line.strip() == "READY", f"cpu worker failed to start: {line!r}"
    t0 = time.perf_counter()
    for p in procs:
        p.stdin.write("go\n")
        p.stdin.flush()
    recs = [json.loads(p.stdout.readline()) for p in procs]
    wall = time.perf_counter() - t0
    for p in procs:
        p.wait()
    agent_steps = sum(r["agents"] * r["steps"] for r in recs)
    fw = sum(r["forwards"] for r in recs)
    sample = (f"{n_proc} processes x 1 thread, one scene (64 agents x 256 polylines) each, first {steps} step(s) of the "
              f"episode, {fw} full-window B=1 forwards in total (the reference pads every window to 32 steps, so its "
              f"cost per step does not depend on t)")
    return agent_steps / wall, n_proc, sample, {"wall_s": wall, "forwards": fw, "s_per_forward_1thread": sum(r["wall"] for r in recs) / max(fw, 1)}


def gpu_torch_run(cfg, weights, dev, groups_per_scene, agents_per_scene, reps=6):
    """Stock torch ops on the same B200 at B=1, two full forwards per focal group per step - the reference's own GPU
    cost structure (policies/autoregressive_policy.py:183-210).  Forward time only (tokenisation, the .cpu() syncs and
    the CPU simulator of the reference are NOT included), so this is an upper bound for that path."""
    import numpy as np
    import torch
    from oracle.model_port import ModelPort
    w = cfg.dataset.waymo
    A, T, P = w.max_num_agents, w.train_context_length, w.max_num_road_polylines
    port = ModelPort(cfg, weights).to(dev)
    rng = np.random.default_rng(0)
    data = {"agent_states": np.concatenate([rng.normal(size=(1, A, T, 7)), np.ones((1, A, T, 1))], -1),
            "agent_types": np.eye(5)[rng.integers(0, 5, (1, A))], "goals": rng.normal(size=(1, A, 5)),
            "actions": rng.integers(0, 1000, (1, A, T)), "rtgs": rng.integers(0, 350, (1, A, T, 3)),
            "timesteps": np.tile(np.arange(T)[None, None, :, None], (1, A, 1, 1)),
            "road_points": np.concatenate([rng.normal(size=(1, P, 100, 2)), np.ones((1, P, 100, 1))], -1),
            "road_types": np.eye(8)[rng.integers(0, 8, (1, P))]}
    data = {k: torch.from_numpy(v).to(dev) for k, v in data.items()}
    for _ in range(3):
        port.forward(data)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = port.forward(data)
        out["action_preds"][0, 0, -1, 0].item()  # the reference reads logits back every forward
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    v = agents_per_scene / (groups_per_scene * 2 * ms * 1e-3)
    return {"value": v, "unit": "agent-steps/s", "ms_per_forward": ms, "kind": "port (stock torch ops, fp32, TF32 off)",
            "sample": f"{reps} B=1 forwards of one {A * T * 3}-token group on cuda; value = {agents_per_scene} agents / "
                      f"({groups_per_scene:.2f} groups per scene x 2 forwards x ms_per_forward); network time only"}


def plan_segments(K, W, steps=90):
    """-> list of timed segments (t_first, n_steps, warm_from) whose phase mix is the episode's (see module docstring).
    warm_from: episode step the untimed run-up to the segment must at least start from (W steps earlier when possible)."""
    q, r = divmod(K, steps)
    n_full = int(round(r * (steps - T_WIN) / steps))
    n_cached = r - n_full
    segs = []
    if n_full:
        t0 = min(T_WIN + W, steps - n_full)
        segs.append((t0, n_full))
    if n_cached:
        segs.append((T_WIN - n_cached, n_cached))
    for _ in range(q):
        segs.append((0, steps))
    return segs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=90)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scenes", type=int, default=256, help="scenes per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunk", type=int, default=256, help="focal groups per workspace chunk")
    ap.add_argument("--wide", action="store_true", help="wide model variant: A=64 agents / P=256 polylines per group")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-sub-batch", type=int, default=64, help="scenes per pipelined sub-batch of the e2e leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-torch-gpu", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=1)
    ap.add_argument("--cpu-worker", type=int, default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()

    if args.cpu_worker is not None:
        cpu_worker(args.cpu_worker, args.cpu_steps, args.wide)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    from ctrlsim_b200.config import default_config
    from ctrlsim_b200.weights import make_weights
    cfg = default_config(wide=args.wide)
    weights = make_weights(cfg, seed=0)
    n_ep = cfg.nocturne.steps

    if args.impl == "reference":
        if rank != 0:
            return
        # bounded sample of the same workload: one scene per host core, min(K, 2) steps each (every step of the
        # reference costs the same: two full padded-window forwards per focal group); ms_per_step is the MEASURED wall
        # time per step of that sample, not an extrapolation to the 256-scene batch
        steps_run = max(1, min(args.steps, 2))
        t0 = time.perf_counter()
        v, cores, sample, det = cpu_reference_run(steps_run, wide=args.wide)
        line = {"impl": "reference", "metric": "agent-steps/s", "value": v, "unit": "agent-steps/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * det["wall_s"] / steps_run,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args.scenes, args.wide), "steps_run": steps_run,
                           "note": "bounded sample: one scene of the workload per host core, the first steps_run steps "
                                   "of the episode; ms_per_step is the measured wall time per step of that sample (the "
                                   "CPU path does not use the GPUs; value is the same for every N)"},
                "cpu_baseline": {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample, **det},
                "e2e": {"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "wall_s_total": time.perf_counter() - t0}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    from ctrlsim_b200 import lib as L
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.model import DeviceModel

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback on the product path)"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    scenes, ids = make_scenes(args.scenes, rank, world)
    model = DeviceModel(cfg, weights, dev)
    lib = model.lib
    pol = B200Policy(cfg, "synthetic", model, seed=0, chunk_groups=args.chunk)
    ev = B200PolicyEvaluator(cfg, pol, scenes=scenes, scene_ids=ids)
    ev.rank, ev.world = 0, 1  # every rank owns all of ITS scenes (already sharded above)
    batch = ev.build_batch(eval_threshold=64)
    ev.rank, ev.world = rank, world
    n_agents = batch.n_evaluated()

    state = {"t": 0}

    def one_step():
        t = state["t"]
        if t == 0:
            pol.reset(batch)
        pol.update_state(batch, t)
        g = pol.predict(batch, t)
        pol.act(batch, t)
        state["t"] = (t + 1) % n_ep
        return g

    def goto(t):  # untimed fast-forward of the running episode to step t (restarting it when t lies behind)
        while state["t"] != t:
            one_step()

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    import ctypes
    segs = plan_segments(args.steps, args.warmup, n_ep)
    ms_total, groups, launches = 0.0, 0, 0
    phase_ms = {"cached": [0.0, 0], "full_window": [0.0, 0]}
    pc_chunks = [0, 0]

    def prefix_stats():
        a, b = ctypes.c_int64(0), ctypes.c_int64(0)
        lib.ctrlsim_prefix_cache_stats(model.handle, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value
    prof = [0.0] * 12
    spans = []
    for t_first, n in segs:
        # run-up: W untimed warm-up steps of the same phase directly in front of the segment (wrapping into the
        # previous episode for a segment that starts at t = 0)
        goto((t_first - args.warmup) % n_ep if t_first >= args.warmup or state["t"] > t_first else 0)
        goto(t_first)
        fence()
        lib.ctrlsim_profile_enable(1)
        l0 = lib.ctrlsim_launch_count()
        pc0 = prefix_stats()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        r0 = clocks.mark()
        torch.cuda.synchronize()
        evs[0].record()
        for i in range(n):
            groups += one_step()
            evs[i + 1].record()
        torch.cuda.synchronize()
        fence()
        spans.append((r0, clocks.mark()))
        ms_total += evs[0].elapsed_time(evs[n])
        launches += lib.ctrlsim_launch_count() - l0
        pc1 = prefix_stats()
        if t_first % n_ep < T_WIN:  # chunks of the cached phase that ran incrementally / had to recompute their window
            pc_chunks[0] += pc1[0] - pc0[0]
            pc_chunks[1] += pc1[1] - pc0[1]
        for i in range(n):
            ph = phase_ms["cached" if (t_first + i) % n_ep < T_WIN else "full_window"]
            ph[0] += evs[i].elapsed_time(evs[i + 1])
            ph[1] += 1
        buf = (ctypes.c_double * 12)()
        lib.ctrlsim_profile_read(buf)
        lib.ctrlsim_profile_enable(0)
        prof = [a + b for a, b in zip(prof, buf)]
    clk = clocks.stop(spans) if rank == 0 else None
    ms = ms_total

    tt = torch.tensor([ms, float(n_agents * args.steps), float(launches), float(groups)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        ms_max, agent_steps, launches_all, groups_all = mx[0].item(), tt[1].item(), tt[2].item(), tt[3].item()
    else:
        ms_max, agent_steps, launches_all, groups_all = tt.tolist()
    value = agent_steps / (ms_max / 1000.0)

    # ---- e2e: public API from host arrays, one full evaluation ----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        fence()
        host_bytes = sum(v.numel() * v.element_size() for k, v in batch.t.items())
        t0 = time.perf_counter()
        ev2 = B200PolicyEvaluator(cfg, pol, scenes=scenes, scene_ids=ids)
        ev2.scenes_presharded = True  # every rank owns all of ITS scenes (sharded above); the all-reduce still spans ranks
        # the call a user makes: host parse + H2D (next sub-batch on a side stream while the current one rolls out) +
        # 90 steps + metrics kernel + D2H of the per-vehicle traces, per sub-batch of 64 scenes; then the one all-reduce
        m_e2e, _ = ev2.evaluate_policy(sub_batch_scenes=args.e2e_sub_batch, eval_threshold=64, keep_traces=True)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        summ = ev2.last_summary
        d2h = sum(v.nbytes for tr in ev2.traces for v in tr.values()) + summ.nbytes
        w = torch.tensor([wall], dtype=torch.float64, device=dev)
        a = torch.tensor([float(ev2.n_evaluated * n_ep)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
            dist.all_reduce(a, op=dist.ReduceOp.SUM)
        e2e = {"value": a.item() / w.item(), "unit": "agent-steps/s",
               "h2d_bytes_per_step": int(host_bytes / n_ep), "d2h_bytes_per_step": int(d2h / n_ep),
               "episode_wall_s": w.item(), "sub_batch_scenes": args.e2e_sub_batch,
               "what": "B200PolicyEvaluator.evaluate_policy(sub_batch_scenes=...) from host scene dicts: whole 90-step "
                       "episodes, the next sub-batch parsed + uploaded while the current one rolls out",
               "metrics": m_e2e if rank == 0 else None}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = _peaks()
    traffic = _ncu_traffic()
    gemm_ms, gemm_fl, gemm_n = prof[0], prof[1], prof[2]
    pool_ms, pool_by, pool_n = prof[3], prof[4], prof[5]
    sa_ms, sa_fl, sa_n = prof[6], prof[7], prof[8]
    ca_ms, ca_fl, ca_n = prof[9], prof[10], prof[11]
    gemm_tf = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    pool_gbs = pool_by / (pool_ms * 1e-3) / 1e9 if pool_ms > 0 else 0.0
    n_full = phase_ms["full_window"][1]
    # "encoder-attn" (BASELINE.json): SURVEY 8(d)(i) asks for both figures.  (1) the product's kernel, fused from the raw
    # points (map_encoder.cu): 1.2 KB in + 8 KB out per polyline, FP32-pipe bound, timed live in the step;
    enc_fused = {"kernel": "map_encode_pool_kernel (product path: point MLP layer 1 -> scores -> softmax -> pooled hidden vectors, "
                           "fused from raw points; W3 / value / output projections folded into the next GEMM)",
                 "bound": "fp32 pipe (not HBM: 0.44 MB per focal group where the unfused chain moved 82 MB)",
                 "unit": "GB/s", "launches": int(pool_n)}
    if pool_n > 0:
        enc_fused.update({"achieved": pool_gbs, "share_of_step": pool_ms / ms, "avg_launch_ms": pool_ms / pool_n,
                          "algorithmic_bytes_per_launch": pool_by / pool_n,
                          # per (point, channel): hidden layer twice (scores pass + pooling pass: 3 FMA + LN affine + ReLU each),
                          # 8 score FMAs, 8 pooling FMAs ~ 54 flops
                          "fp32_gflops_per_launch": pool_by / pool_n / (NPTS * 12 + 1 + 8192) * NPTS * 256 * 54 / 1e9})
        enc_fused["fp32_tflops"] = enc_fused["fp32_gflops_per_launch"] / enc_fused["avg_launch_ms"]
    else:
        enc_fused.update({"achieved": None, "note": "not exercised: no full-window step in the timed region"})
    # (2) the HBM-streaming pooling attention over precomputed [points, 256] features (map_pool_kernel: one TMA bulk copy
    # per 100 KB polyline tile, exported as ctrlsim_map_pool) - the kernel the roofline target is quoted on - timed here
    # on one chunk's worth of polylines (5 GB of features, far beyond L2) with CUDA events on the launching stream
    enc = {"kernel": "map_pool_kernel (HBM-streaming polyline pooling attention over precomputed features, TMA bulk + mbarrier "
                     "ring; the unfused variant, not launched by the product path any more)", "bound": "hbm",
           "peak": pk["hbm"], "unit": "GB/s", "peak_source": pk["hbm_src"], "traffic": traffic.get("map_pool_kernel"),
           "timed": "in isolation inside this bench.py run, same polyline count as one chunk of the step"}
    try:
        n_poly = args.chunk * 200
        feats = torch.randn(n_poly, 100, 256, device=dev)
        pv = (torch.rand(n_poly, 100, device=dev) > 0.1).to(torch.uint8)
        okp = torch.ones(n_poly, dtype=torch.uint8, device=dev)
        U = torch.randn(8, 256, device=dev) * 0.1
        outp = torch.empty(n_poly, 8, 256, device=dev)
        st_ = torch.cuda.current_stream(dev).cuda_stream
        for _ in range(3):
            lib.ctrlsim_map_pool(feats.data_ptr(), pv.data_ptr(), okp.data_ptr(), U.data_ptr(), outp.data_ptr(), n_poly, st_)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(5):
            lib.ctrlsim_map_pool(feats.data_ptr(), pv.data_ptr(), okp.data_ptr(), U.data_ptr(), outp.data_ptr(), n_poly, st_)
        eb.record()
        torch.cuda.synchronize()
        t_ms = ea.elapsed_time(eb) / 5
        by = n_poly * (100 * 256 * 4 + 100 + 8 * 256 * 4)
        enc.update({"achieved": by / t_ms / 1e6, "frac": by / t_ms / 1e6 / pk["hbm"], "launches": 5, "avg_launch_ms": t_ms,
                    "algorithmic_bytes_per_launch": by})
        del feats, pv, okp, outp
    except Exception as e:  # noqa: BLE001 - the headline must not die on this side measurement
        enc.update({"achieved": None, "frac": None, "note": f"side measurement failed: {e}"})
    line = {
        "metric": "agent-steps/s", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.scenes, args.wide), "scenes_per_gpu": args.scenes,
                   "controlled_agents_rank0": n_agents,
                   "focal_groups_per_step_avg": groups_all / args.steps / world, "chunk_groups": args.chunk,
                   "timed_segments": [{"t_first": a, "steps": b} for a, b in segs],
                   "l2": "per-step working set (>10 GB of activations per chunk) far exceeds the 126 MB L2; no explicit flush",
                   "weights": "random-init (deterministic generator), reference architecture",
                   "map_cache": "per-focal polyline-encoder cache for steps 0..31 (ctrlsim_attach_map_cache): " + ("on" if pol.use_map_cache else "off"),
                   "simulator": "FreeCar + Box2D vehicle-vehicle contact response: " + ("off" if os.environ.get("CTRLSIM_CONTACTS", "1") == "0" else "on")},
        "phases": {"cached_ms_per_step": phase_ms["cached"][0] / max(phase_ms["cached"][1], 1), "cached_steps": phase_ms["cached"][1],
                   "cached_chunks_incremental": pc_chunks[0], "cached_chunks_recomputed": pc_chunks[1],
                   "full_window_ms_per_step": phase_ms["full_window"][0] / max(n_full, 1), "full_window_steps": n_full,
                   "episode_mix": "58 full-window : 32 cached steps per 90-step episode"},
        "gpu_launches": int(launches_all),
        "clocks": clk,
        "roofline": {"kernel": "gemm_tc_ta_kernel (every nn.Linear: tcgen05 kind::tf32 with the activation operand in TMEM, 3-product hi/lo split = fp32-accurate, 3 tensor flops per counted flop)", "bound": "tensor", "achieved": gemm_tf,
                     "peak": pk["tf"] / 1.0, "unit": "TFLOP/s", "frac": gemm_tf / pk["tf"],
                     "traffic": traffic.get("gemm_tc_ta_kernel", traffic.get("gemm_tc_tma_kernel")),
                     # the chip runs this kernel at its 1 kW cap (profiles/r02r_power_probe.txt): cuBLAS tf32 sustains
                     # 596 TFLOP/s on the same box, so no 3-product tf32 GEMM can sustain more than 199 counted
                     "power_limited": {"cublas_tf32_sustained_tflops": 596.3, "issued_tf32_tflops": 3.0 * gemm_tf,
                                       "frac_of_tf32_x3_ceiling": 3.0 * gemm_tf / 596.3,
                                       "source": "profiles/r02r_power_probe.txt (tools/power_probe.py, this pool's B200)"},
                     "peak_source": pk["tf_src"], "share_of_step": gemm_ms / ms, "launches": int(gemm_n),
                     "avg_launch_ms": gemm_ms / max(gemm_n, 1), "algorithmic_flops_per_launch": gemm_fl / max(gemm_n, 1)},
        "encoder_attn": enc,
        "encoder_attn_fused": enc_fused,
        "kernel_shares": {"gemm": gemm_ms / ms, "map_pool": pool_ms / ms, "decoder_self_attn": sa_ms / ms,
                          "decoder_cross_attn": ca_ms / ms,
                          "decoder_self_attn_tflops": sa_fl / (sa_ms * 1e-3) / 1e12 if sa_ms > 0 else 0.0,
                          "decoder_cross_attn_tflops": ca_fl / (ca_ms * 1e-3) / 1e12 if ca_ms > 0 else 0.0},
    }
    if e2e:
        line["e2e"] = e2e
    if world == 1:  # the baselines are reported by the single-GPU run only
        if not args.no_torch_gpu:
            line["gpu_torch_baseline"] = gpu_torch_run(cfg, weights, dev, groups_all / args.steps / args.scenes,
                                                       n_agents / args.scenes)
        if not args.no_cpu:
            del model, pol, ev, batch
            v, cores, sample, det = cpu_reference_run(args.cpu_steps, wide=args.wide)
            line["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample, **det}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
