"""N>1 host path on CPU: scenes shard over ranks, the only collective is the all-reduce of the summary buffer.
Runs two gloo processes on 127.0.0.1 (no GPU): sharding + evaluated-set draw reproducibility + summary additivity."""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden

NAMES = ("plumbing", "crowded", "sparse")


def summary_from_record(cfg, g):
    """Host restatement of the per-scene part of ctrlsim_metrics' output layout (8 scalars + 8x200 histograms)."""
    from oracle.policy_port import MetricsPort
    mp_ = MetricsPort(cfg)
    rec = {k: g[k] for k in ("existence", "reward", "pos", "gt_pos", "vel", "gt_speed", "heading", "gt_heading",
                             "gt_accel", "accel", "gt_nearest_dist", "nearest_dist")}
    mp_.add_scene(rec, [int(v) for v in g["evaluated"]])
    w = cfg.dataset.waymo
    cat = {k: np.concatenate(v) for k, v in mp_.samples.items()}
    hist = np.zeros((8, 200))
    e_lin, e_ang, e_nd = np.arange(201) * 0.5 * (100 / 30), np.arange(201) * 0.5 - 50, np.arange(201) * 0.5 * (100 / 40)
    hist[0] = np.histogram(np.clip(cat["lin_sim"], 0, 30), bins=e_lin)[0]
    hist[1] = np.histogram(np.clip(cat["lin_gt"], 0, 30), bins=e_lin)[0]
    hist[2] = np.histogram(np.clip(cat["ang_sim"], -50, 50), bins=e_ang)[0]
    hist[3] = np.histogram(np.clip(cat["ang_gt"], -50, 50), bins=e_ang)[0]
    gq = (np.clip(cat["acc_gt"], w.min_accel, w.max_accel) - w.min_accel) / (w.max_accel - w.min_accel)
    gq = np.round(gq * 19) / 19 * 20 - 10
    e_acc = np.arange(21) * 2 - 20
    hist[4, :20] = np.histogram(cat["acc_sim"], bins=e_acc)[0]
    hist[5, :20] = np.histogram(gq, bins=e_acc)[0]
    hist[6] = np.histogram(np.clip(cat["nd_sim"], 0, 40), bins=e_nd)[0]
    hist[7] = np.histogram(np.clip(cat["nd_gt"], 0, 40), bins=e_nd)[0]
    head = np.array([sum(mp_.goal), len(mp_.goal), sum(mp_.coll), sum(mp_.off), len(mp_.coll), sum(mp_.ade), sum(mp_.fde), 0.0])
    return np.concatenate([head, hist.flatten()])


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ctrlsim_b200.config import default_config
    from ctrlsim_b200.evaluator import B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    cfg = default_config()
    # (1) sharding: scene k -> rank k mod world, global scene ids kept, evaluated sets drawn by ONE seeded stream
    scenes = [make_scene(40 + i, n_vehicles=12, n_roads=2, n_chunks=2) for i in range(5)]
    stub = types.SimpleNamespace(model=types.SimpleNamespace(device="cpu"))
    ev = B200PolicyEvaluator(cfg, stub, scenes=scenes, scene_ids=[100 + i for i in range(5)])
    b = ev.build_batch(eval_threshold=4)
    mine = [100 + i for i in range(5) if i % world == rank]
    assert b.t["scene_id"].tolist() == mine
    ev1 = B200PolicyEvaluator(cfg, stub, scenes=scenes, scene_ids=[100 + i for i in range(5)])
    ev1.rank, ev1.world = 0, 1
    full = ev1.build_batch(eval_threshold=4)
    assert [full.evaluated_ids[i] for i in range(5) if i % world == rank] == b.evaluated_ids
    # (2) the one collective: summaries are additive over ranks
    parts = [summary_from_record(cfg, load_golden(n)[0]) for i, n in enumerate(NAMES) if i % world == rank]
    s = torch.from_numpy(np.sum(parts, axis=0))
    dist.all_reduce(s)
    if rank == 0:
        np.save(out, s.numpy())
    # (3) planner-vs-adversary: scenes k -> rank k mod world, per-scene statistics exchanged with all_gather_object
    from conftest import load_planner_adversary_golden
    from ctrlsim_b200.planner_adversary import B200PlannerAdversaryEvaluator, CatAdversary, PlannerAdversaryStats
    recs, spec, ref = load_planner_adversary_golden("policies")
    pa_scenes = [{"json": {"objects": []}, "preproc": {}, "tag": k} for k in range(len(recs))]
    pa = B200PlannerAdversaryEvaluator(cfg, stub, CatAdversary(), scenes=pa_scenes, pairs=spec["pairs"],
                                       adv_trajs=[r["adv_pos"] for r in recs])
    assert [sc["tag"] for sc in pa.scenes] == [k for k in range(len(recs)) if k % world == rank]
    st = PlannerAdversaryStats(cfg)
    for sc, (ego, adv) in zip(pa.scenes, pa.pairs):
        st.add_scene(recs[sc["tag"]], ego, adv)
    got = pa.gather(st).compute()
    for k, v in ref.items():
        assert (np.isnan(v) and np.isnan(got[k])) or abs(got[k] - v) < 1e-9 * max(1.0, abs(v)), (k, got[k], v)
    # (4) fewer scenes than ranks: the rank without a scene builds an empty batch, skips the rollout (its policy stub
    # has no verbs: any call would raise) and still joins the one all-reduce with a zero summary
    ev0 = B200PolicyEvaluator(cfg, stub, scenes=scenes[:1], scene_ids=[100])
    b0 = ev0.build_batch(eval_threshold=4)
    assert b0.S == (1 if rank == 0 else 0)
    if rank == 0:
        s1 = torch.from_numpy(summary_from_record(cfg, load_golden("sparse")[0]))
        dist.all_reduce(s1)
    else:
        from ctrlsim_b200.evaluator import B200Policy
        stub.reset = lambda b: B200Policy.reset(stub, b)
        stub.update_state = lambda b, t: B200Policy.update_state(stub, b, t)
        stub.predict = lambda b, t: B200Policy.predict(stub, b, t)
        stub.act = lambda b, t: B200Policy.act(stub, b, t)
        ev0.rollout(b0)
        assert b0.n_evaluated() == 0 and b0.contact_overflow() == 0
        s1 = torch.from_numpy(ev0.summarize(b0))
        want = summary_from_record(cfg, load_golden("sparse")[0])
        assert np.array_equal(s1.numpy(), want)  # zeros + rank 0's part
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_summary(tmp_path, cfg):
    out = str(tmp_path / "summary.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    from ctrlsim_b200.evaluator import B200PolicyEvaluator
    from oracle.policy_port import MetricsPort
    got = B200PolicyEvaluator.metrics_from_summary(np.load(out))
    ref = MetricsPort(cfg)
    for n in NAMES:
        g = load_golden(n)[0]
        rec = {k: g[k] for k in ("existence", "reward", "pos", "gt_pos", "vel", "gt_speed", "heading", "gt_heading",
                                 "gt_accel", "accel", "gt_nearest_dist", "nearest_dist")}
        ref.add_scene(rec, [int(v) for v in g["evaluated"]])
    want = ref.compute()
    for k, v in want.items():
        assert abs(got[k] - v) < 1e-9, (k, got[k], v)
