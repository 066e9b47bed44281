"""CPU tests (no GPU): the oracle is pinned against the reference's own known answers and against fixtures produced by
the unmodified reference (tests/golden/, oracle/make_golden.py); plus host logic and the C-ABI surface."""
import ctypes
import math
import os
import sys
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden

needs_reference = pytest.mark.skipif(
    not os.path.exists(os.path.join(ROOT, "oracle", "_ref")) or not os.path.isdir("/root/reference"),
    reason="needs the reference sources / build (build container only)")


# ------------------------------------------------------------------------------------------------ sampler / RNG
def test_philox_known_answer_vectors():
    """Random123 KAT for philox4x32-10."""
    from ctrlsim_b200.philox import philox4x32
    z = philox4x32(np.zeros(4, np.uint32), np.zeros(2, np.uint32))
    assert [hex(int(v)) for v in z] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = philox4x32(np.full(4, 0xFFFFFFFF, np.uint32), np.full(2, 0xFFFFFFFF, np.uint32))
    assert [hex(int(v)) for v in f] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_sampler_is_a_faithful_multinomial():
    from oracle import sampler
    x = np.array([0.0, 1.0, 2.0, -1.0, 0.5], np.float32)
    w = sampler.weights_from_x(x).astype(np.float64)
    p = np.exp(x.astype(np.float64) - 2.0)
    assert np.abs(w / w.sum() - p / p.sum()).max() < 1e-6
    d = np.linspace(-80, 0, 4001).astype(np.float32)
    assert np.abs(sampler.exp_spec(d) / np.exp(d.astype(np.float64)) - 1).max() < 3e-7
    draws = np.array([sampler.sample_from_x(x, 3, 0, a, 0, 3) for a in range(4000)])
    freq = np.bincount(draws, minlength=5) / 4000.0
    assert np.abs(freq - p / p.sum()).max() < 0.03
    # tilt: the reference adds tilt * linspace(0,1,350) in float64 (dataset.py:342-348)
    lg = np.zeros(350, np.float32)
    assert sampler.rtg_x(lg, 10.0)[-1] == np.float32(10.0) and sampler.rtg_x(lg, 10.0)[0] == 0


def test_nucleus_keep_matches_the_reference_rule():
    """oracle.sampler.nucleus_keep (integer weights) against the reference's own torch statement on float
    probabilities (policies/autoregressive_policy.py:216-230): identical kept sets on generic inputs; hand cases for
    ties, p = 0 (arg-max only) and p >= 1 (everything with a non-zero weight)."""
    import torch
    from oracle import sampler
    rng = np.random.default_rng(5)
    p = 0.8
    for r in range(60):
        x = (rng.standard_normal(1000) * rng.uniform(0.5, 6.0)).astype(np.float32)
        probs = torch.softmax(torch.from_numpy(x), dim=0)
        sp, si = torch.sort(probs, descending=True)
        sel = torch.cumsum(sp, dim=-1) < p
        sel = torch.cat([sel.new_ones(1), sel[:-1]])
        ref_keep = np.zeros(1000, bool)
        ref_keep[si[sel].numpy()] = True
        got = sampler.nucleus_keep(sampler.weights_from_x(x), p)
        assert (got == ref_keep).all(), (r, int(got.sum()), int(ref_keep.sum()))
    w = np.array([4, 4, 8, 1, 4, 0], np.uint64)  # sorted: 8 | 4(i0) 4(i1) 4(i4) | 1 | 0 ; total 21
    assert sampler.nucleus_keep(w, 0.0).tolist() == [False, False, True, False, False, False]
    assert sampler.nucleus_keep(w, 0.5).tolist() == [True, False, True, False, False, False]      # 8 < 10.5 -> +4(i0): 12
    assert sampler.nucleus_keep(w, 0.6).tolist() == [True, True, True, False, False, False]       # 12 < 12.6 -> +4(i1)
    assert sampler.nucleus_keep(w, 1.0).tolist() == [True, True, True, True, True, False]         # 20 < 21 -> +1; 21 !< 21
    assert sampler.nucleus_keep(w, 1.5).all()
    # draws only ever land in the kept set
    x = (rng.standard_normal(1000) * 3).astype(np.float32)
    keep = sampler.nucleus_keep(sampler.weights_from_x(x), 0.8)
    draws = {sampler.sample_from_x_nucleus(x, 0.8, 1, 2, a, 0, 3) for a in range(300)}
    assert all(keep[d] for d in draws) and len(draws) > 1


# ------------------------------------------------------------------------------------------------ simulator oracle
def test_geometry_known_answers_c_oracle():
    """nocturne/cpp/tests/src/geometry/polygon_test.cc:60-86 and intersection_test.cc:52-76 on the C restatement."""
    from oracle import sim_port as sp
    eps = 1e-5
    sq = [(0, 0), (1, 0), (1, 1), (0, 1)]
    tri = [(1, 2), (2, 1), (2, 2)]
    assert not sp.poly_intersects(sq, tri) and not sp.poly_intersects(tri, sq)
    tri = [(1 - eps, 1 - eps), (2, 0), (2, 2)]
    assert sp.poly_intersects(sq, tri) and sp.poly_intersects(tri, sq)
    tri = [(1, 1), (2, 0), (2, 2)]
    assert sp.poly_intersects(sq, tri) and sp.poly_intersects(tri, sq)
    dia = [(1, 0), (0, 1), (-1, 0), (0, -1)]
    for s0, s1, want in [((0, 0.5), (0, -0.5), True), ((-0.5, -0.5), (-0.5, -1.0), True), ((-1, 0.5), (1, 1), True),
                         ((1, 1), (-1, -1), True), ((-1, -1 - eps), (1, -1 - eps), False), ((-3, 0.5), (-2, 1), False)]:
        assert sp.poly_segment_intersects(dia, s0, s1) is want


def test_sim_oracle_known_answer_from_reference():
    """output of the real nocturne_cpp on a 2-vehicle scene (recorded from the reference)."""
    from oracle import sim_port as sp
    T = 91

    def obj(x, y, hdeg, speed):
        vx, vy = speed * math.cos(math.radians(hdeg)), speed * math.sin(math.radians(hdeg))
        return {"type": "vehicle", "length": 4.5, "width": 2.0, "position": [{"x": x, "y": y}] * T, "heading": [hdeg] * T,
                "velocity": [{"x": vx, "y": vy}] * T, "valid": [True] * T, "goalPosition": {"x": x, "y": y}}
    scen = {"objects": [obj(1000.0, 2000.0, 0.0, 10.0), obj(1030.0, 2000.0, 180.0, 5.0)],
            "roads": [{"type": "road_edge", "geometry": [{"x": 990.0 + i, "y": 2005.0} for i in range(100)]}]}
    sim = sp.ScenePort(sp.parse_scenario(scen))
    want = {0: (1001.0175, 2000.0510, 0.02269, 10.1874, 0), 9: (1010.9034, 2001.7715, 0.24659, 11.9577, 0),
            14: (1016.7851, 2003.9357, 0.38627, 12.9344, 1), 19: (1022.7221, 2007.1527, 0.53677, 13.9050, 0)}
    for k in range(20):
        sim.set_action(0, 2.0, 0.1)
        sim.set_action(1, -1.0, 0.0)
        sim.step(0.1)
        if k in want:
            x, y, h, s, edge = want[k]
            assert abs(sim.position()[0, 0] - x) < 2e-4 and abs(sim.position()[0, 1] - y) < 2e-4
            assert abs(sim.heading()[0] - h) < 1e-5 and abs(sim.speed()[0] - s) < 1e-4
            assert bool(sim.collisions()[1][0]) == bool(edge)
    assert abs(sim.speed()[1] - 3.0) < 1e-5 and abs(sim.heading()[1] + np.pi) < 1e-6


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref")) or not os.path.isdir("/root/reference"),
                    reason="needs the reference build (build container only)")
def test_sim_oracle_bit_exact_vs_reference_nocturne(tmp_path, cfg):
    """Random actions on a contact-free scene: the C restatement is bit-identical to the real nocturne_cpp."""
    import json
    from ctrlsim_b200.synth import make_scene
    from oracle import ref_shims, sim_port
    ref_shims.install()
    import nocturne
    sc = make_scene(5, n_vehicles=6, n_roads=6, n_chunks=4, lane_ids=[3])
    path = tmp_path / "s.json"
    path.write_text(json.dumps(sc["json"]))
    sim = nocturne.Simulation(scenario_path=str(path), config=cfg.nocturne["scenario"])
    vehs = sim.getScenario().vehicles()
    for v in vehs:
        v.expert_control = False
        v.physics_simulated = True
    port = sim_port.ScenePort(sim_port.parse_scenario(sc["json"]))
    rng = np.random.default_rng(0)
    for t in range(60):
        p = np.array([[v.getPosition().x, v.getPosition().y] for v in vehs])
        assert (p == port.position()).all() and (np.array([v.getHeading() for v in vehs]) == port.heading()).all()
        assert (np.array([v.getSpeed() for v in vehs]) == port.speed()).all()
        assert (np.array([int(v.collision_type_edge) == 2 for v in vehs]) == port.collisions()[1]).all()
        for i, v in enumerate(vehs):
            a, s = rng.uniform(-3, 3), rng.uniform(-0.3, 0.3)
            if t % 7 == 3:
                a = 0.0
            if a > 0:
                v.acceleration = a
            else:
                v.brake(abs(a))
            v.steering = s
            port.set_action(i, a, s)
        sim.step(0.1)
        port.step(0.1)


# ------------------------------------------------------------------------------------------------ model oracle
def test_mask_rule_matches_its_definition():
    from oracle.model_port import causal_mask_rule
    A = 3
    m = causal_mask_rule(A, 2, 3).numpy()

    def idx(t, a, k):
        return (t * A + a) * 3 + k
    assert m[idx(1, 0, 0), idx(0, 2, 2)]          # everything in the past is visible
    assert m[idx(1, 0, 0), idx(1, 2, 0)]          # all current states are visible, even of later agents
    assert not m[idx(1, 0, 0), idx(1, 0, 1)]      # a state token does not see its own rtg
    assert m[idx(1, 0, 1), idx(1, 0, 1)] and not m[idx(1, 0, 1), idx(1, 0, 2)]
    assert m[idx(1, 0, 2), idx(1, 0, 1)] and not m[idx(1, 0, 2), idx(1, 1, 1)]
    assert not m[idx(0, 1, 2), idx(1, 0, 0)]      # never the future


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference tree (build container only)")
def test_mask_rule_equals_reference_get_causal_mask(cfg):
    import sys
    import types
    from oracle.model_port import causal_mask_rule
    if "utils" not in sys.modules:
        mod = types.ModuleType("utils")
        mod.__path__ = ["/root/reference/utils"]
        sys.modules["utils"] = mod
    from utils.train_utils import get_causal_mask
    small = cfg.copy()
    small.dataset.waymo.max_num_agents = 4
    ref = get_causal_mask(small, 5, 3)
    assert ((ref == 0).numpy() == causal_mask_rule(4, 5, 3).numpy()).all()
    import copy
    dt = copy.deepcopy(small)  # decision-transformer token order (rtg, state, action): the state token sits at index 1
    dt.model.decision_transformer = True
    assert ((get_causal_mask(dt, 5, 3) == 0).numpy() == causal_mask_rule(4, 5, 3, state_index=1).numpy()).all()


@pytest.mark.parametrize("name,step", [("plumbing", 9), ("crowded", 33), ("sparse", 89)])
def test_model_port_matches_reference_logits(cfg, name, step):
    """oracle/model_port.py vs logits the reference CtRLSim.forward produced on the same tokens (fixtures)."""
    from ctrlsim_b200.weights import make_weights
    from oracle.model_port import ModelPort
    g, spec, _ = load_golden(name)
    model = ModelPort(cfg, make_weights(cfg, **spec["weights"]))
    t = step
    ti = t if t < 32 else 31
    data = {k: torch.from_numpy(g[f"in_{t}_{k}"][None]) for k in ("agent_states", "agent_types", "goals", "actions",
                                                                   "timesteps", "road_points", "road_types")}
    data["rtgs"] = torch.from_numpy(g[f"in_{t}_rtgs_pass1"][None])
    out = model.forward(data)
    assert np.abs(out["rtg_preds"][0, :, ti].numpy() - g[f"rtg_logits_{t}_0"]).max() < 5e-5
    data["rtgs"] = torch.from_numpy(g[f"in_{t}_rtgs_pass2"][None])
    out = model.forward(data)
    assert np.abs(out["action_preds"][0, :, ti].numpy() - g[f"action_logits_{t}_0"]).max() < 5e-5


# ------------------------------------------------------------------------------------------------ rollout oracle
def test_rollout_port_matches_reference_prefix(cfg):
    """BASELINE config 1 (1 scene, 8 vehicles, 32 polylines, 10 steps): oracle port vs the unmodified reference
    evaluator - sampled bins identical, trajectories identical, groups identical."""
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    from oracle.model_port import ModelPort
    from oracle.policy_port import RolloutPort
    g, spec, _ = load_golden("plumbing")
    steps = 10
    port = RolloutPort(cfg, ModelPort(cfg, make_weights(cfg, **spec["weights"])), seed=0, tilts=tuple(spec["tilts"]),
                       eval_threshold=64)
    sc = make_scene(**spec["scene"])
    rec = port.run_scene(0, sc["json"], sc["preproc"], max_steps=steps)
    assert (rec["rtg_idx"][:steps] == g["rtg_idx"][:steps]).all()
    assert (rec["act_idx"][:steps] == g["act_idx"][:steps]).all()
    for k in ("pos", "vel", "heading", "existence", "accel", "steer", "reward", "nearest_dist", "gt_nearest_dist"):
        assert np.abs(rec[k][:, :steps] - g[k][:, :steps]).max() < 1e-9, k
    for t in range(steps):
        for gi, d in enumerate(rec["groups"][t]):
            assert d["focal"] == g["group_focal"][t, gi] and (d["members"] == g["group_members"][t, gi]).all()
            assert d["served"] == [int(v) for v in g["group_served"][t, gi] if v >= 0]


def test_rollout_port_dt_real_time_rewards_matches_reference_prefix():
    """the decision-transformer baseline as cfgs/policy/dt.yaml runs it (RTGs tracked in real time from
    the dense reward, one forward per focal group) - oracle port vs the unmodified reference evaluator
    (tests/golden/rollout_dt.npz): sampled action bins, trajectories, the dense reward and the RTG series."""
    from ctrlsim_b200.config import dt_config
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    from oracle.model_port import ModelPort
    from oracle.policy_port import RolloutPort
    g, spec, _ = load_golden("dt")
    cfg = dt_config()
    steps = 8
    port = RolloutPort(cfg, ModelPort(cfg, make_weights(cfg, **spec["weights"])), seed=0, eval_threshold=64,
                       predict_rtgs=False, discretize_rtgs=False, real_time_rewards=True, max_return=True)
    sc = make_scene(**spec["scene"])
    rec = port.run_scene(0, sc["json"], sc["preproc"], max_steps=steps, logit_steps=(0,))
    assert (rec["act_idx"][:steps] == g["act_idx"][:steps]).all() and (rec["rtg_idx"][:steps] == -1).all()
    for k in ("pos", "vel", "heading", "existence", "accel", "steer", "reward", "nearest_dist", "gt_nearest_dist",
              "dense_reward", "rtgs"):
        assert np.abs(rec[k][:, :steps] - g[k][:, :steps]).max() < 1e-9, k
    assert (g["rtgs"][:, 0] == (10.0, 90.0, 90.0)).all() and np.abs(np.diff(g["rtgs"], axis=1)).max() > 0.1
    for gi in range(len(rec["groups"][0])):
        assert np.abs(rec["logits"][(0, gi)]["action_logits"] - g[f"action_logits_0_{gi}"]).max() < 5e-5


def test_dense_reward_port_reproduces_the_reference_episode():
    """The dense reward and the RTG bookkeeping of all 91 steps of the reference's DT episode, recomputed by the oracle
    from the recorded states: bit-identical.  Pins two things no unit test of the functions shows: road-edge points are
    the simulator's float32 copies, and the goal / collision terms are those of STEP 0 (evaluator.py:112-113,136-138
    index the reward history with 0)."""
    from ctrlsim_b200.config import dt_config
    from ctrlsim_b200.synth import make_scene
    from oracle import dense_reward_port as drp
    g, spec, _ = load_golden("dt")
    w = dt_config().dataset.waymo
    sc = make_scene(**spec["scene"])
    edges = drp.road_edge_polylines(sc["json"])
    n = g["pos"].shape[0]
    tr = drp.RtgTracker(drp.initial_rtgs(w, sc["preproc"], n), [int(v) for v in g["evaluated"]], max_return=True)
    assert g["reward"][:, 1:, 7].max() == 1.0 and g["reward"][:, 0, 7].max() == 0.0  # an off-road flag appears later only
    for t in range(91):
        dense, _ = drp.dense_reward_step(w, g["pos"][:, t], g["existence"][:, t], g["reward"][:, 0], edges)
        assert (dense == g["dense_reward"][:, t]).all(), t
        if t < 90:
            rtg = tr.rtg[0] if t == 0 else tr.advance(g["dense_reward"][:, t - 1])
            assert (rtg == g["rtgs"][:, t]).all(), t


@needs_reference
def test_initial_rtgs_port_matches_the_reference_dataset(tmp_path, cfg):
    """RTGs at t = 0 without max_return / min_return: reverse cumulative sum of compute_rewards over the logged episode of
    the *_physics.pkl (datasets/rl_waymo/dataset_ctrl_sim.py:38-92 eval branch) - oracle restatement and the product's
    loader (ctrlsim_b200.scenario.initial_rtgs) against the reference dataset object."""
    import pickle
    from oracle import ref_shims
    from oracle.dense_reward_port import initial_rtgs
    from ctrlsim_b200.scenario import initial_rtgs as product_initial_rtgs
    ref_shims.install()
    from datasets.rl_waymo import RLWaymoDatasetCtRLSim
    rng = np.random.default_rng(3)
    n = 9
    pre = {"idx": 0, "num_agents": n, "road_points": np.zeros((4, 100, 3)), "road_types": np.zeros((4, 8)),
           "ag_data": rng.normal(size=(n, 90, 8)), "ag_actions": np.zeros((n, 90, 2)), "ag_types": np.zeros((n, 5)),
           "last_exist_timesteps": np.zeros(n), "ag_rewards": rng.uniform(0, 1, size=(n, 90, 8)),
           "veh_edge_dist_rewards": rng.normal(size=(n, 90)) * 0.2, "veh_veh_dist_rewards": rng.uniform(0, 1, size=(n, 90)),
           "filtered_ag_ids": list(range(n)), "ag_goals": np.zeros((n, 90, 5))}
    pre["ag_data"][:, :, -1] = (rng.uniform(size=(n, 90)) > 0.2).astype(np.float64)
    pre["ag_rewards"][:, :, [0, 6, 7]] = (pre["ag_rewards"][:, :, [0, 6, 7]] > 0.8).astype(np.float64)
    dset = RLWaymoDatasetCtRLSim.__new__(RLWaymoDatasetCtRLSim)
    dset.cfg_dataset, dset.preprocess, dset.mode = cfg.dataset.waymo, True, "eval"
    for k, v in (("POS_TARGET_ACHIEVED_REW_IDX", 0), ("HEADING_TARGET_ACHIEVED_REW_IDX", 1), ("SPEED_TARGET_ACHIEVED_REW_IDX", 2),
                 ("POS_GOAL_SHAPED_REW_IDX", 3), ("SPEED_GOAL_SHAPED_REW_IDX", 4), ("HEADING_GOAL_SHAPED_REW_IDX", 5),
                 ("VEH_VEH_COLLISION_REW_IDX", 6), ("VEH_EDGE_COLLISION_REW_IDX", 7)):
        assert getattr(dset, k) == v
    rtgs, _, _ = dset.get_data(pre, 0)
    want = np.concatenate([rtgs[:, 0, :1], rtgs[:, 0, 3:]], -1)
    assert np.abs(initial_rtgs(cfg.dataset.waymo, pre, n) - want).max() < 1e-12
    assert np.abs(product_initial_rtgs(cfg, pre) - want).max() < 1e-12


def test_metrics_port_matches_reference_metrics(cfg):
    """S7: feeding the reference's own recorded trajectories through the oracle metrics reproduces its metrics dict."""
    from oracle.policy_port import MetricsPort
    for name in ("plumbing", "crowded", "sparse", "dt"):
        g, spec, ref = load_golden(name)
        mp = MetricsPort(cfg)
        rec = {k: g[k] for k in ("existence", "reward", "pos", "gt_pos", "vel", "gt_speed", "heading", "gt_heading",
                                 "gt_accel", "accel", "gt_nearest_dist", "nearest_dist")}
        mp.add_scene(rec, [int(v) for v in g["evaluated"]])
        got = mp.compute()
        for k, v in ref.items():
            assert abs(got[k] - v) < 1e-9, (name, k, got[k], v)


def test_tokenizer_port_matches_reference_inputs(cfg):
    """T2/T3 teacher-forced on the reference's recorded history: the port's group tokens at sliding-window steps equal
    the tensors the reference fed its model (fixture in_<t>_*)."""
    from ctrlsim_b200.synth import make_scene
    from oracle.policy_port import RolloutPort
    g, spec, _ = load_golden("sparse")
    sc = make_scene(**spec["scene"])
    n = g["pos"].shape[0]
    port = RolloutPort(cfg, model=None, eval_threshold=64)
    for t in (9, 33, 89):
        ep = {"states": np.zeros((n, 90, 8)), "actions": np.zeros((n, 90, 2)), "rtgs": np.zeros((n, 90, 3)),
              "goals": np.zeros((n, 90, 5)), "timesteps": np.zeros((n, 90, 1)), "types": np.tile(np.eye(5)[1], (n, 1)),
              "road_points": sc["preproc"]["road_points"], "road_types": sc["preproc"]["road_types"]}
        ep["states"][:, :t + 1, :2], ep["states"][:, :t + 1, 2:4] = g["pos"][:, :t + 1], g["vel"][:, :t + 1]
        ep["states"][:, :t + 1, 4], ep["states"][:, :t + 1, 7] = g["heading"][:, :t + 1], g["existence"][:, :t + 1]
        ep["states"][:, :t + 1, 5:7] = g["size"][:, None]
        ep["actions"][:, :t, 0], ep["actions"][:, :t, 1] = g["accel"][:, :t], g["steer"][:, :t]
        ep["rtgs"][:, :t] = g["rtgs"][:, :t]
        ep["timesteps"][:, :t + 1, 0] = np.arange(t + 1)
        gl = g["goal"]
        ep["goals"][:] = np.stack([gl[:, 0], gl[:, 1], gl[:, 3] * np.cos(gl[:, 2]), gl[:, 3] * np.sin(gl[:, 2]),
                                   gl[:, 2]], -1)[:, None]
        focal = int(g["group_focal"][t, 0])
        closest = np.array([v for v in g["group_members"][t, 0] if v >= 0])
        tok = port.tokenize(ep, t, focal, closest)
        for k in ("agent_states", "agent_types", "goals", "actions", "road_points", "road_types"):
            assert np.abs(tok[k].astype(np.float32) - g[f"in_{t}_{k}"].astype(np.float32)).max() < 1e-5, (t, k)
        assert (tok["rtgs"] == g[f"in_{t}_rtgs_pass1"]).all()
        assert (tok["timesteps"] == g[f"in_{t}_timesteps"]).all()


# ------------------------------------------------------------------------------------------------ C-ABI surface
def test_library_loads_and_exports_every_declared_symbol():
    from ctrlsim_b200 import lib
    from ctrlsim_b200.build import build_all
    header = open(os.path.join(ROOT, "include", "ctrlsim_b200.h")).read()
    declared = set(re.findall(r"\b(ctrlsim_[a-z_0-9]+)\s*\(", header))
    assert declared, "no prototypes parsed from the header"
    assert set(lib.EXPORTS) == declared
    paths = build_all()  # the reference-default geometry and the wide one (-DCTRLSIM_WIDE), same ABI
    assert len(paths) == 2
    for path in paths:
        so = ctypes.CDLL(path)
        for name in declared:
            assert hasattr(so, name), f"{name} declared in include/ctrlsim_b200.h but not exported by {path}"
        so.ctrlsim_abi_version.restype = ctypes.c_int
        assert so.ctrlsim_abi_version() == lib.ABI_VERSION


def test_batch_struct_matches_header_field_order():
    from ctrlsim_b200 import lib
    header = open(os.path.join(ROOT, "include", "ctrlsim_b200.h")).read()
    body = header[header.index("typedef struct CtrlSimBatch {"):header.index("} CtrlSimBatch;")]
    fields = re.findall(r"\*\s*([a-z_0-9]+);", body)
    assert fields == [n for n, _, _ in lib.BATCH_FIELDS]


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ctrlsim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_c_abi_fails_loudly_without_a_gpu_and_on_bad_arguments(cfg):
    """The product path has no CPU fallback: on a box without a CUDA device ctrlsim_create returns a negative status
    and an explanatory message; argument errors are reported the same way (no exceptions cross the C boundary)."""
    import ctypes as C
    import torch
    from ctrlsim_b200 import lib as L
    so = L.load(build_if_missing=False)
    # size queries are pure host arithmetic and must agree with the documented layouts
    per_block = 200 * 256 * 4 + 200
    assert so.ctrlsim_map_cache_bytes(3, 64) == 3 * 64 * per_block + 256
    slot = so.ctrlsim_prefix_cache_bytes(8, 1)
    assert slot >= 8 * (4 * 2304 * 512 * 4 + 4 * 224 * 512 * 4 + 224) and so.ctrlsim_prefix_cache_bytes(8, 5) == 5 * slot
    assert so.ctrlsim_workspace_bytes(None, 2) > so.ctrlsim_workspace_bytes(None, 1) > 0
    # bad arguments
    h = C.c_void_p()
    assert so.ctrlsim_create(None, C.byref(h)) < 0 and b"null" in so.ctrlsim_last_error()
    cc = L.make_config(cfg)
    cc.abi_version = 1  # stale ABI
    assert so.ctrlsim_create(C.byref(cc), C.byref(h)) < 0 and b"ABI" in so.ctrlsim_last_error()
    cc = L.make_config(cfg)
    cc.hidden_dim = 128  # a geometry the kernels are not specialised to
    assert so.ctrlsim_create(C.byref(cc), C.byref(h)) < 0 and b"specialised" in so.ctrlsim_last_error()
    if not torch.cuda.is_available():
        cc = L.make_config(cfg)
        rc = so.ctrlsim_create(C.byref(cc), C.byref(h))
        assert rc < 0, "ctrlsim_create must not succeed without a CUDA device"
        msg = so.ctrlsim_last_error().decode()
        assert "CUDA" in msg or "sm_100" in msg, msg


def _parse_scenario_scalar(scen, steps=90, moving_threshold=0.2, speed_threshold=0.05):
    """Element-by-element statement of the loader rules (scenario.cc:893-1057, utils/sim.py:20-65) the vectorised
    ctrlsim_b200.scenario.parse_scenario must reproduce bit for bit."""
    import math
    from ctrlsim_b200.scenario import _normalize_angle_f32
    objs = [o for o in scen["objects"] if bool(o["valid"][0]) and o["type"] == "vehicle"]
    n, T1 = len(objs), steps + 1
    gt = np.zeros((n, T1, 4), np.float32)
    gt_valid = np.zeros((n, T1), np.uint8)
    target = np.zeros((n, 4), np.float32)
    moving = np.zeros(n, bool)
    for i, o in enumerate(objs):
        gp = o.get("goalPosition", {"x": 0.0, "y": 0.0})
        target[i, :2] = (gp["x"], gp["y"])
        for t in range(len(o["position"])):
            x, y = np.float32(o["position"][t]["x"]), np.float32(o["position"][t]["y"])
            h = _normalize_angle_f32(o["heading"][t])
            vx, vy = np.float32(o["velocity"][t]["x"]), np.float32(o["velocity"][t]["y"])
            sp = np.sqrt(np.float32(vx * vx + vy * vy))
            if t < T1:
                gt[i, t] = (x, y, h, sp)
                gt_valid[i, t] = x != np.float32(-10000.0)
            if bool(o["valid"][t]):
                target[i, 2], target[i, 3] = h, sp
                dx, dy = x - target[i, 0], y - target[i, 1]
                dist = np.sqrt(np.float32(dx * dx + dy * dy))
                if sp > np.float32(speed_threshold) or dist > np.float32(moving_threshold):
                    moving[i] = True
    segs = []
    for road in scen["roads"]:
        g = road["geometry"]
        if road["type"] != "road_edge" or isinstance(g, dict):
            continue
        for k in range(len(g) - 1):
            segs.append((g[k]["x"], g[k]["y"], g[k + 1]["x"], g[k + 1]["y"]))
    return dict(gt=gt, gt_valid=gt_valid, moving=moving, target=target, segs=np.asarray(segs, np.float32).reshape(-1, 4))


def test_vectorised_scenario_parser_is_bit_identical_to_the_scalar_rules():
    from ctrlsim_b200.scenario import parse_scenario
    from ctrlsim_b200.synth import make_scene
    cases = [make_scene(0, n_vehicles=12, n_roads=2, n_chunks=3),
             make_scene(3, n_vehicles=9, n_roads=1, n_chunks=2, frac_short=0.5, frac_parked=0.3)]
    # a track with odd headings (wrap-around on both sides) and an early disappearance
    o = cases[1]["json"]["objects"][0]
    for t in range(len(o["heading"])):
        if o["valid"][t]:
            o["heading"][t] = -540.0 + 13.7 * t
    for sc in cases:
        a = _parse_scenario_scalar(sc["json"])
        b = parse_scenario(sc["json"], 90, 0.2, 0.05)
        assert np.array_equal(a["gt"], b["gt"]) and a["gt"].dtype == b["gt"].dtype
        assert np.array_equal(a["gt_valid"], b["gt_valid"]) and np.array_equal(a["moving"], b["moving"])
        assert np.array_equal(a["segs"], b["segs"]) and a["segs"].dtype == b["segs"].dtype
        # goal heading / speed of vehicles that never disappear = the last valid state
        keep = b["gt_valid"].all(1)
        assert np.array_equal(a["target"][keep, 2:].astype(np.float64), b["goal"][keep, 2:])


# ---- on-disk formats: checkpoint import and the evaluator's file reader ------------------------------
def test_param_spec_matches_the_reference_state_dict(cfg):
    """tests/golden/state_dict_spec.json = names and shapes of the unmodified reference CtRLSim().state_dict()
    (oracle/make_state_dict_spec.py): the checkpoint importer must expect exactly these tensors."""
    import json
    from ctrlsim_b200.weights import param_spec
    with open(os.path.join(ROOT, "tests", "golden", "state_dict_spec.json")) as f:
        ref = json.load(f)
    mine = {k: list(s) for k, s, _ in param_spec(cfg)}
    assert mine == ref
    assert sum(int(np.prod(s)) for s in ref.values()) == 8285762  # 8.29 M parameters


def test_lightning_checkpoint_import_round_trip_and_errors(tmp_path, cfg):
    import torch
    from ctrlsim_b200.checkpoint import CheckpointError, check_state_dict, load_checkpoint, save_checkpoint
    from ctrlsim_b200.weights import make_weights
    w = make_weights(cfg, seed=3)
    path = str(tmp_path / "model.ckpt")
    save_checkpoint(w, path, cfg)
    blob = torch.load(path, map_location="cpu", weights_only=False)
    assert set(blob) >= {"state_dict", "pytorch-lightning_version", "hyper_parameters"}  # what load_from_checkpoint reads
    got = load_checkpoint(path, cfg)
    assert list(got) == list(w)
    assert all(got[k].dtype == np.float32 and np.array_equal(got[k], w[k]) for k in w)
    # wrapper prefixes (EMA / DataParallel / compiled modules) and a bare state_dict file
    torch.save({"model." + k: torch.from_numpy(v) for k, v in w.items()}, path)
    got = load_checkpoint(path, cfg)
    assert all(np.array_equal(got[k], w[k]) for k in w)
    # the unused future-state head may be absent; anything else missing or mis-shaped is named in the error
    slim = {k: v for k, v in w.items() if not k.startswith("decoder.predict_future_states")}
    assert len(check_state_dict(slim, cfg)) == len(slim) < len(w)
    broken = dict(w)
    del broken["encoder.map_encoder.map_seeds"]
    broken["decoder.predict_action.mlp.3.weight"] = broken["decoder.predict_action.mlp.3.weight"][:999]
    with pytest.raises(CheckpointError) as e:
        check_state_dict(broken, cfg)
    assert "encoder.map_encoder.map_seeds" in str(e.value) and "(999, 256)" in str(e.value)
    with pytest.raises(CheckpointError):
        check_state_dict(dict(w, stray=np.zeros(3, np.float32)), cfg, strict_unexpected=True)
    # arbitrary pickled objects (an omegaconf tree in the reference's checkpoints) are refused unless the caller vouches
    # for the file; a missing file is reported as such, not retried
    import fractions
    torch.save({"state_dict": {k: torch.from_numpy(v) for k, v in w.items()}, "hyper_parameters": fractions.Fraction(1, 3)}, path)
    with pytest.raises(CheckpointError, match="trusted=True"):
        load_checkpoint(path, cfg)
    got = load_checkpoint(path, cfg, trusted=True)
    assert all(np.array_equal(got[k], w[k]) for k in w)
    with pytest.raises(FileNotFoundError):
        load_checkpoint(str(tmp_path / "absent.ckpt"), cfg, trusted=True)


def test_file_reader_counts_only_evaluated_scenes_like_the_reference(tmp_path, cfg):
    """evaluate_policy stops after num_files_to_evaluate // partitions EVALUATED scenes; scenes without a *_physics.pkl
    or without a candidate agent are skipped and do not count (policy_evaluator.py:436-437,445-446,461-464,492)."""
    import copy
    import types
    from ctrlsim_b200.evaluator import B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene, write_dataset
    scenes = [make_scene(70 + i, n_vehicles=6, n_roads=2, n_chunks=2) for i in range(6)]
    for o in scenes[1]["json"]["objects"]:  # scene 1: nobody moves -> no candidate agent
        o["position"] = [dict(o["position"][0]) for _ in o["position"]]
        o["velocity"] = [{"x": 0.0, "y": 0.0} for _ in o["velocity"]]
        o["goalPosition"] = dict(o["position"][0])
    paths = write_dataset(str(tmp_path), scenes)
    os.remove(os.path.join(paths["preprocess_dir"], "test", scenes[2]["name"] + "_physics.pkl"))  # scene 2: no preproc
    c = copy.deepcopy(cfg)
    c.dataset_root, c.nocturne_waymo_val_folder = paths["dataset_root"], paths["nocturne_waymo_val_folder"]
    c.eval.num_files_to_evaluate, c.eval.partitions = 6, 2  # -> 3 evaluated scenes
    stub = types.SimpleNamespace(model=types.SimpleNamespace(device="cpu"))
    ev = B200PolicyEvaluator(c, stub)
    mine, ids, parsed, evs, thr = ev.select_scenes()
    assert ids == [0, 3, 4]  # file indices: 1 and 2 skipped, 5 beyond the limit
    assert [s["name"] for s in mine] == [scenes[i]["name"] + ".json" for i in ids]
    assert all(len(e) > 0 for e in evs) and thr == c.eval.multi_agent_eval_threshold
    # the same draw as an in-memory evaluation of the accepted scenes (one seeded stream over all scenes walked)
    ev_mem = B200PolicyEvaluator(c, stub, scenes=[scenes[i] for i in (0, 1, 3, 4)], scene_ids=[0, 1, 3, 4])
    _, ids_mem, _, evs_mem, _ = ev_mem.select_scenes()
    assert ids_mem == ids and evs_mem == evs


# ---- planner vs adversary --------------------------------------------------------------------------
def _pa_setup(cfg, spec):
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    scenes = [make_scene(**s) for s in spec["scenes"]]
    return scenes, make_weights(cfg, **spec["weights"])


def _same_metrics(got, ref, tol=1e-9):
    assert set(got) == set(ref)
    for k, v in ref.items():
        if isinstance(v, float) and np.isnan(v):
            assert np.isnan(got[k]), (k, got[k])
        else:
            assert abs(got[k] - v) <= tol * max(1.0, abs(v)), (k, got[k], v)


@pytest.mark.parametrize("name", ["policies", "cat"])
def test_planner_adversary_metrics_match_the_reference(cfg, name):
    """Both restatements of update_running_statistics / compute_metrics (planner_adversary_evaluator.py:201-428) - the
    oracle's and the product's host-side PlannerAdversaryStats - on the reference's own per-scene records."""
    from conftest import load_planner_adversary_golden
    from ctrlsim_b200.planner_adversary import PlannerAdversaryStats
    from oracle.planner_adversary_port import PlannerAdversaryMetricsPort
    recs, spec, ref = load_planner_adversary_golden(name)
    port, stats = PlannerAdversaryMetricsPort(cfg), PlannerAdversaryStats(cfg)
    for rec in recs:
        ego, adv = (int(x) for x in rec["ego_adv"])
        port.add_scene(rec, ego, adv)
        stats.add_scene(rec, ego, adv)
    _same_metrics(port.compute(), ref)
    _same_metrics(stats.compute(), ref)


def test_planner_adversary_port_matches_reference_prefix(cfg):
    """Two policies on one world (own RTG series, context sets and sampler seeds): first steps of the oracle port vs
    the unmodified reference evaluator - sampled bins bit-exact, trajectories within float tolerance."""
    from conftest import load_planner_adversary_golden
    from oracle.model_port import ModelPort
    from oracle.planner_adversary_port import PlannerAdversaryPort
    from oracle.policy_port import RolloutPort
    recs, spec, _ = load_planner_adversary_golden("policies")
    scenes, weights = _pa_setup(cfg, spec)
    mp = ModelPort(cfg, weights)
    port = PlannerAdversaryPort(cfg, RolloutPort(cfg, mp, seed=spec["seeds"][0], tilts=spec["tilts_planner"]),
                                RolloutPort(cfg, mp, seed=spec["seeds"][1], tilts=spec["tilts_adversary"]))
    steps, k = 11, 1  # crosses t = 9, where the policies take over from log replay
    g, sc, (ego, adv) = recs[k], scenes[k], spec["pairs"][k]
    rec = port.run_scene(k, sc["json"], sc["preproc"], ego, adv, max_steps=steps)
    for role in ("planner", "adversary"):
        assert (rec[role]["act_idx"][:steps] == g[f"{role}_act_idx"][:steps]).all()
        assert (rec[role]["rtg_idx"][:steps] == g[f"{role}_rtg_idx"][:steps]).all()
        assert np.abs(rec[role]["rtgs"][:, :steps] - g[f"{role}_rtgs"][:, :steps]).max() < 1e-12
    assert (rec["planner"]["act_idx"][:steps, ego] >= 0).all() and (rec["adversary"]["act_idx"][:steps, adv] >= 0).all()
    assert np.abs(rec["pos"][:, :steps] - g["pos"][:, :steps]).max() < 1e-3
    assert np.abs(rec["accel"][:, :steps] - g["accel"][:, :steps]).max() < 1e-6


def test_scripted_adversary_port_matches_reference_prefix(cfg):
    """CAT adversary: log replay until step 8, then the scripted trajectory through the inverse bicycle model."""
    from conftest import load_planner_adversary_golden
    from oracle.model_port import ModelPort
    from oracle.planner_adversary_port import PlannerAdversaryPort
    from oracle.policy_port import RolloutPort
    recs, spec, _ = load_planner_adversary_golden("cat")
    scenes, weights = _pa_setup(cfg, spec)
    port = PlannerAdversaryPort(cfg, RolloutPort(cfg, ModelPort(cfg, weights), seed=spec["seeds"][0],
                                                 tilts=spec["tilts_planner"]), None)
    steps = 14
    g, sc, (ego, adv) = recs[0], scenes[0], spec["pairs"][0]
    rec = port.run_scene(0, sc["json"], sc["preproc"], ego, adv, adv_pos=g["adv_pos"], max_steps=steps)
    assert (rec["planner"]["act_idx"][:steps] == g["planner_act_idx"][:steps]).all()
    assert np.abs(rec["pos"][:, :steps] - g["pos"][:, :steps]).max() < 1e-3
    assert np.abs(rec["accel"][:, :steps] - g["accel"][:, :steps]).max() < 1e-6
    assert np.abs(rec["steer"][:, :steps] - g["steer"][:, :steps]).max() < 1e-6
    # the scripted track really differs from the logged one after the hand-over
    assert np.abs(g["accel"][adv, 9:steps]).max() > 1e-2


# ---- Box2D contact response (SURVEY 7.3-1 stage 2) --------------------------------------------------------------------
@needs_reference
@pytest.mark.parametrize("mode", ["rear", "side", "head", "pile", "dense"])
def test_contact_solver_bit_exact_vs_reference_nocturne(mode):
    """Vehicles that hit each other: the C restatement of Box2D's broad phase bookkeeping, polygon manifold, warm-started
    sequential-impulse solver (block solver, 8 + 3 iterations) and island sleeping is bit-identical to the real
    nocturne_cpp / Box2D for the whole run - through the impact, the pushing phase, a vehicle vanishing mid-contact
    and the separation; "dense" = 64 vehicles with random controls, up to 15 touching contacts at once.  Each case runs
    in a fresh process: the reference's b2World is a process singleton and a Simulation dropped without reset() leaves
    its bodies in it, where they collide with the next case at the same coordinates (tests/contact_case.py)."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "contact_case.py"), mode], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bit-exact" in r.stdout, r.stdout[-500:]


@pytest.mark.parametrize("name", ["plumbing", "crowded", "sparse"])
def test_sim_oracle_with_contacts_reproduces_reference_episodes(name):
    """The controls the unmodified reference evaluator applied (tests/golden/rollout_*.npz) pushed through the C
    simulator restatement WITH contacts give the reference's trajectories bit for bit over all 90 steps - including
    'plumbing' (vehicles touch from step 12 on) and 'crowded' (30 vehicles, contacts from step 16 on), where the
    contact-free subset is metres off right after the first contact."""
    from ctrlsim_b200.synth import make_scene
    from oracle import sim_port
    g, spec, _ = load_golden(name)
    parsed = sim_port.parse_scenario(make_scene(**spec["scene"])["json"])
    gt = sim_port.ground_truth(parsed, 90)
    port = sim_port.ScenePort(parsed, contacts=True)
    ev = set(int(v) for v in g["evaluated"])
    for t in range(91):
        ex = g["existence"][:, t].astype(bool)
        assert (port.position()[ex] == g["pos"][:, t][ex]).all(), (name, t)
        assert (port.heading()[ex] == g["heading"][:, t][ex]).all(), (name, t)
        assert (port.collisions()[0][ex] == (g["reward"][:, t, 6][ex] == 1)).all(), (name, t)
        if t == 90:
            break
        for i in range(parsed["n"]):
            if t >= 9 and i in ev:  # policy.act (autoregressive_policy.py:256-274)
                if not g["existence"][i, t]:
                    port.teleport(i, -1000000, -1000000)
            else:                   # apply_gt_action (evaluators/evaluator.py:160-193)
                exists = gt[i, t, 4] and gt[i, t + 1, 4]
                if t > 0 and g["existence"][i, t] == 0:
                    exists = 0
                if not exists:
                    port.teleport(i, -1000000, -1000000)
            port.set_action(i, g["accel"][i, t], g["steer"][i, t])
        port.step(0.1)


@pytest.mark.parametrize("name", ["plumbing", "crowded", "sparse"])
def test_gpu_acceptance_logic_on_cpu_emulation_of_gpu_arithmetic(cfg, name):
    """The GPU simulator is the oracle's algorithm with one arithmetic difference: sinf / cosf / tanf are evaluated in
    fp64 and rounded once, glibc's are within 1 ulp but not always correctly rounded.  The oracle built with the GPU's
    trig (oracle/_build/libsim_oracle_fp64trig.so) therefore predicts the GPU trace on the CPU: the reference's controls
    are pushed through it and the result goes through the very acceptance function the GPU parity test uses
    (tests/parity_checks.py).  Measured on the B200, the GPU numbers equal this emulation digit for digit
    (crowded: 7.32421875e-4 m, 2.3126602e-5 rad)."""
    from ctrlsim_b200.synth import make_scene
    from oracle.policy_port import RolloutPort
    from parity_checks import check_rollout_vs_reference
    g, spec, _ = load_golden(name)
    sc = make_scene(**spec["scene"])
    port = RolloutPort(cfg, None, contacts=True, fp64_trig=True)
    ctx = port.setup_scene(0, sc["json"])
    n, sim, rec = ctx["n"], ctx["sim"], ctx["rec"]
    evaluated = [int(v) for v in g["evaluated"]]
    next_act = np.zeros((n, 2))
    for t in range(90):
        port.observe(ctx, t)
        next_act[:, 0], next_act[:, 1] = g["accel"][:, t], g["steer"][:, t]
        port.apply_controls(ctx, t, evaluated, next_act)
        sim.step(0.1)
    port.observe(ctx, 90)
    tr = {"tr_pos": rec["pos"][None], "tr_vel": rec["vel"][None], "tr_heading": rec["heading"][None],
          "tr_exist": rec["existence"][None], "tr_action": np.stack([rec["accel"], rec["steer"]], -1)[None],
          "tr_reward": rec["reward"][None], "tr_nearest": np.stack([rec["nearest_dist"], rec["gt_nearest_dist"]], -1)[None],
          "tr_rtg_idx": g["rtg_idx"].transpose(1, 0, 2)[None], "tr_act_idx": g["act_idx"].T[None]}
    assert check_rollout_vs_reference(tr, g, name, trig="fp64") == 90
    ex = g["existence"].astype(bool)
    dpos = np.abs(rec["pos"] - g["pos"])[ex].max()
    dhead = np.abs(rec["heading"] - g["heading"])[ex].max()
    if name == "crowded":  # the figures the B200 run reported for the same episode
        assert abs(dpos - 0.000732421875) < 1e-12 and abs(dhead - 2.3126602172851562e-05) < 1e-12, (dpos, dhead)
    if name == "sparse":
        assert dpos == 0.0


def test_glibc_trig_restatement_equals_this_machines_libm(tmp_path):
    """ctrlsim_b200/csrc/glibc_trig.h (the sinf / cosf the GPU simulator uses under CTRLSIM_TRIG=glibc) compiled for the
    host - same source as the device build - against sinf / cosf of this machine's libm over 6e6 arguments in
    [-8, 8], [-0.9, 0.9] and [-119, 119], and tanf on [-pi/4, pi/4]: no difference.  (1.2e8 arguments: none; 'evaluate in fp64 and round once'
    differs in 1.3 % of them, tools/trig_check.cpp.)"""
    import subprocess
    exe = str(tmp_path / "trig_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", os.path.join(ROOT, "tools", "trig_check.cpp"), "-o", exe])
    r = subprocess.run([exe, "3000000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "differs from libm in 0 sinf / 0 cosf" in r.stdout, r.stdout[-600:]
    assert "tanf on [-pi/4, pi/4]: port differs from libm in 0 of" in r.stdout, r.stdout[-600:]


@pytest.mark.parametrize("name", ["plumbing", "crowded"])
def test_glibc_trig_mode_makes_the_contact_episodes_bit_exact(cfg, name):
    """CPU emulation of the GPU's CTRLSIM_TRIG=glibc mode (oracle built on the product's glibc_trig.h instead of libm):
    the reference's controls reproduce the reference's trajectories bit for bit over all 90 steps - also 'crowded',
    where the default mode (fp64 trig) drifts by 0.73 mm while vehicles push each other."""
    from ctrlsim_b200.synth import make_scene
    from oracle.policy_port import RolloutPort
    g, spec, _ = load_golden(name)
    sc = make_scene(**spec["scene"])
    port = RolloutPort(cfg, None, contacts=True, fp64_trig="glibc_port")
    ctx = port.setup_scene(0, sc["json"])
    n, sim, rec = ctx["n"], ctx["sim"], ctx["rec"]
    evaluated = [int(v) for v in g["evaluated"]]
    next_act = np.zeros((n, 2))
    for t in range(90):
        port.observe(ctx, t)
        next_act[:, 0], next_act[:, 1] = g["accel"][:, t], g["steer"][:, t]
        port.apply_controls(ctx, t, evaluated, next_act)
        sim.step(0.1)
    port.observe(ctx, 90)
    ex = g["existence"].astype(bool)
    assert (rec["pos"][ex] == g["pos"][ex]).all() and (rec["heading"][ex] == g["heading"][ex]).all()
    assert (rec["reward"][:, :, 6][ex] == g["reward"][:, :, 6][ex]).all() and g["reward"][:, :, 6][ex].any()
    # ... and passes the (strict, default-mode) acceptance function of the GPU parity test
    from parity_checks import check_rollout_vs_reference
    tr = {"tr_pos": rec["pos"][None], "tr_vel": rec["vel"][None], "tr_heading": rec["heading"][None],
          "tr_exist": rec["existence"][None], "tr_action": np.stack([rec["accel"], rec["steer"]], -1)[None],
          "tr_reward": rec["reward"][None], "tr_nearest": np.stack([rec["nearest_dist"], rec["gt_nearest_dist"]], -1)[None],
          "tr_rtg_idx": g["rtg_idx"].transpose(1, 0, 2)[None], "tr_act_idx": g["act_idx"].T[None]}
    assert check_rollout_vs_reference(tr, g, name, trig="glibc") == 90


@pytest.mark.parametrize("mode", ["rear", "side", "head", "pile", "dense"])
def test_product_contact_code_equals_oracle_on_the_host(tmp_path, mode):
    """The PRODUCT's contact code (ctrlsim_b200/csrc/sim_contacts.cuh - what sim.cu compiles for the GPU), built for the
    host (tests/host_contacts_shim.cpp) and put between the oracle's FreeCar step and its Vehicle::Step / collision
    flags, against the oracle's own world step: bit-identical bodies every step of the collision cases (impact, pushing,
    a vehicle teleported away mid-contact, three-car pile-up).  Keeps the CUDA source checkable without a GPU."""
    import ctypes
    import subprocess
    from contact_case import collision_scene
    from oracle import sim_port
    so = str(tmp_path / "libhc.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-w",
                           os.path.join(ROOT, "tests", "host_contacts_shim.cpp"), "-o", so])
    H = ctypes.CDLL(so)
    F, U8 = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint8)
    H.hc_init.argtypes = [F, F, F, ctypes.c_int, ctypes.c_int, F]
    H.hc_world_step.argtypes = [F, F, F, ctypes.c_int, ctypes.c_int, F, U8, ctypes.c_float]
    H.hc_num_touching.argtypes = [F, ctypes.c_int, ctypes.c_int]
    fields = ("px", "py", "cx", "cy", "lcx", "lcy", "ang", "vx", "vy", "om", "sleep_t", "thr", "brk", "steer", "awake")
    assert H.hc_body_fields() == len(fields) + 1
    if mode == "dense":  # BASELINE config-2 scene: 64 vehicles, random controls -> dozens of pairs, islands of several contacts
        from ctrlsim_b200.synth import make_scene
        parsed = sim_port.parse_scenario(make_scene(2)["json"])
    else:
        parsed = sim_port.parse_scenario(collision_scene(mode)["json"])
    A = sim_port.ScenePort(parsed, contacts=True)   # oracle, its own world step
    B = sim_port.ScenePort(parsed, contacts=False)  # oracle FreeCar / Vehicle::Step around the product's world step
    n = N = parsed["n"]
    L = sim_port.lib()
    fp = lambda a: a.ctypes.data_as(F)

    def pack():
        return np.ascontiguousarray(np.stack([B.arr[k].astype(np.float32) for k in fields] + [np.zeros(n, np.float32)]))

    def unpack(body):
        for k, row in zip(fields, body):
            B.arr[k][:] = row.astype(B.arr[k].dtype)

    cstate = np.zeros(H.hc_words(N), np.float32)
    body = pack()
    H.hc_init(fp(body), fp(B.arr["len"]), fp(B.arr["wid"]), N, n, fp(cstate))
    rng = np.random.default_rng(1)
    touched = 0
    for t in range(90 if mode == "dense" else 45):
        for k in ("px", "py", "ang", "vx", "vy", "om", "sleep_t", "awake", "ox", "oy", "heading", "speed", "coll_veh"):
            assert (A.arr[k] == B.arr[k]).all(), (mode, t, k)
        tele = np.zeros(n, np.uint8)
        for i in range(n):
            a, s = (0.5, 0.0) if (i < 3 and mode != "dense") else (rng.uniform(-3, 3), rng.uniform(-0.3, 0.3))
            if (t == 30 and i == 1) or (mode == "dense" and t >= 60 and i % 7 == 3):
                A.teleport(i, -1000000, -1000000)
                B.teleport(i, -1000000, -1000000)
                tele[i] = 1
            A.set_action(i, a, s)
            B.set_action(i, a, s)
        A.step(0.1)
        L.simo_freecar_all(ctypes.byref(B.s), np.float32(0.1))
        body = pack()
        H.hc_world_step(fp(body), fp(B.arr["len"]), fp(B.arr["wid"]), N, n, fp(cstate), tele.ctypes.data_as(U8), np.float32(0.1))
        unpack(body)
        L.simo_finish_step(ctypes.byref(B.s), sim_port._fp(B.segs), len(B.segs))
        assert H.hc_num_touching(fp(cstate), N, n) == A.n_touching(), (mode, t)
        touched = max(touched, A.n_touching())
    assert touched >= {"pile": 2, "dense": 6}.get(mode, 1), (mode, touched)


# ---- groundwork for SURVEY 8(f) N1: dense real-time reward ------------------------------------------------------------
@needs_reference
def test_dense_reward_port_matches_the_reference_functions(cfg):
    """oracle/dense_reward_port.py (scalar loops) against the reference's own array code: signed distance to road-edge
    polylines (utils/data.py:152-290), nearest-vehicle distance and compute_rewards (datasets/rl_waymo/dataset.py:
    202-275) as evaluators/evaluator.py:106-140 combines them.  Product counterpart: dense_reward_kernel (GPU tests test_dt_*)."""
    from oracle import ref_shims
    from oracle.dense_reward_port import dense_reward_step, signed_distance_to_polylines
    ref_shims.install()
    from utils.data import compute_distance_to_road_edge
    from datasets.rl_waymo import RLWaymoDatasetCtRLSim
    rng = np.random.default_rng(0)
    th = np.linspace(0, 2 * np.pi, 41)
    ring = np.stack([30 * np.cos(th), 20 * np.sin(th)], -1)                      # closed, counter-clockwise
    ring[-1] = ring[0]
    edge = np.stack([np.linspace(-60, 60, 25), 35 + 3 * np.sin(np.linspace(0, 6, 25))], -1)  # open, wavy
    zig = np.array([[-50.0, -40.0], [-20.0, -30.0], [0.0, -45.0], [25.0, -28.0], [55.0, -42.0]])
    polys = [ring, edge, zig]
    pts = rng.uniform(-70, 70, size=(400, 2))
    pts = np.concatenate([pts, ring[::5] * 1.0, edge[::4] + 1e-9, np.array([[30.0, 0.0], [-60.0, 35.0]])])
    ref = compute_distance_to_road_edge(pts[:, 0][None], pts[:, 1][None], polys)
    got = np.array([signed_distance_to_polylines(x, y, polys) for x, y in pts])
    assert np.abs(got - ref).max() < 1e-9, np.abs(got - ref).max()
    assert (np.sign(got[np.abs(got) > 1e-6]) == np.sign(ref[np.abs(got) > 1e-6])).all()
    # one step of compute_dense_reward
    dset = RLWaymoDatasetCtRLSim.__new__(RLWaymoDatasetCtRLSim)
    dset.cfg_dataset = cfg.dataset.waymo
    n = 12
    pos = rng.uniform(-40, 40, size=(n, 2))
    exist = (rng.uniform(size=n) > 0.25).astype(np.float64)
    rew = rng.uniform(0, 1, size=(n, 8))
    rew[:, [0, 1, 2, 6, 7]] = (rew[:, [0, 1, 2, 6, 7]] > 0.7).astype(np.float64)
    proc = rew[:, None, :] * exist[:, None, None]
    ag = pos[:, None, :]
    edge_r = dset.compute_dist_to_nearest_road_edge_rewards(ag, polys) * exist[:, None]
    agx = np.concatenate([pos, exist[:, None]], 1)[:, None, :]
    vv = dset.compute_dist_to_nearest_vehicle_rewards(agx.copy(), normalize=False) * exist[:, None]
    w = cfg.dataset.waymo
    vvn = np.clip(vv, 0.0, w.max_veh_veh_distance) / w.max_veh_veh_distance
    allr = dset.compute_rewards(agx, proc, edge_r, vvn)
    ref_dense = np.concatenate([allr[:, :, :1], allr[:, :, 3:]], -1)[:, 0]
    dense, nearest = dense_reward_step(w, pos, exist, rew, polys)
    assert np.abs(dense - ref_dense).max() < 1e-9, np.abs(dense - ref_dense).max()
    assert np.abs(nearest - vv[:, 0]).max() < 1e-9


@needs_reference
def test_model_port_dt_variant_matches_the_reference_modules(cfg):
    """Decision-transformer baseline (cfgs/model/dt.yaml: continuous RTG inputs, no RTG head): oracle/model_port.py against
    the reference Encoder / Decoder built with that configuration and its own random initialisation.  Oracle groundwork
    for SURVEY 8(f) N1; product counterpart: the decision_transformer mode of the library (GPU test
    test_dt_forward_matches_reference_logits)."""
    import copy
    import torch
    from oracle import ref_shims
    from oracle.model_port import ModelPort
    ref_shims.install()
    from models import CtRLSim
    c = copy.deepcopy(cfg)
    c.model.decision_transformer, c.model.predict_rtg, c.model.predict_future_states = True, False, False
    torch.manual_seed(0)
    ref = CtRLSim(c).eval()
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    assert sd["encoder.embed_rtg_goal.weight"].shape == (c.model.hidden_dim, 1) and "decoder.predict_rtg.mlp.0.weight" not in sd
    A, T, P = c.dataset.waymo.max_num_agents, c.dataset.waymo.train_context_length, c.dataset.waymo.max_num_road_polylines
    g = torch.Generator().manual_seed(1)
    n = 7
    st = torch.zeros(1, A, T, 8, dtype=torch.float64)
    st[0, :n, :, :7] = torch.randn(n, T, 7, generator=g, dtype=torch.float64) * torch.tensor([20, 20, 5, 5, 1, 0.1, 0.1]) + torch.tensor([0, 0, 0, 0, 0, 4.8, 2.0])
    st[0, :n, :, 7] = (torch.rand(n, T, generator=g) > 0.1).double()
    types = -torch.ones(1, A, 5, dtype=torch.float64)
    types[0, :n] = torch.eye(5, dtype=torch.float64)[1]
    rp = torch.zeros(1, P, 100, 3, dtype=torch.float64)
    rp[0, :40, :, :2] = torch.randn(40, 100, 2, generator=g, dtype=torch.float64) * 30
    rp[0, :40, :, 2] = (torch.rand(40, 100, generator=g) > 0.2).double()
    rt = -torch.ones(1, P, 8, dtype=torch.float64)
    rt[0, :40] = torch.eye(8, dtype=torch.float64)[torch.randint(0, 8, (40,), generator=g)]
    data = {"agent_states": st, "agent_types": types, "goals": torch.randn(1, A, 5, generator=g, dtype=torch.float64) * 10,
            "actions": torch.randint(0, 1000, (1, A, T), generator=g).double(),
            "rtgs": torch.rand(1, A, T, 3, generator=g, dtype=torch.float64),
            "timesteps": torch.arange(T)[None, None, :, None].expand(1, A, T, 1).clone(),
            "road_points": rp, "road_types": rt}
    from torch_geometric.data import HeteroData
    md = HeteroData({"agent": {k: data[k] for k in ("agent_states", "agent_types", "goals", "actions", "rtgs", "timesteps")},
                     "map": {"road_points": data["road_points"], "road_types": data["road_types"]}})
    md["agent"]["moving_agent_mask"] = torch.ones(1, A)
    with torch.no_grad():
        want = ref(md, eval=True)["action_preds"]
    got = ModelPort(c, sd).forward(data)
    assert "rtg_preds" not in got
    d = (got["action_preds"][0, :n] - want[0, :n]).abs().max().item()
    assert d < 5e-5, d


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref")) or not os.path.isdir("/root/reference"),
                    reason="needs the reference tree and its built simulator (build container only)")
def test_one_agent_and_two_agent_selection_match_the_reference_functions(tmp_path, cfg):
    """cfg.eval.eval_mode = one_agent / two_agent (cfgs/eval/base.yaml:13-14; one_agent is the reference's default):
    the evaluated vehicles are a random "interesting" pair (or its first vehicle).  The host-side restatement
    (scenario.interesting_pairs + B200PolicyEvaluator.select_scenes) is compared with the UNMODIFIED
    PolicyEvaluator.find_interesting_agent / find_interesting_pair (policy_evaluator.py:308-416) run on the real
    nocturne_cpp objects of the same scenes, draw for draw with the same seeded generator."""
    import random
    import types
    from ctrlsim_b200.evaluator import B200PolicyEvaluator
    from ctrlsim_b200.scenario import interesting_pairs, parse_scenario
    from ctrlsim_b200.synth import make_scene, write_dataset
    from oracle import ref_harness, ref_shims
    ref_shims.install()
    from evaluators import PolicyEvaluator
    from utils.sim import get_ground_truth_states, get_moving_vehicles
    scenes = [make_scene(700 + i, n_vehicles=10 + 6 * i, n_roads=1 + i % 2, n_chunks=3, frac_short=0.4) for i in range(5)]
    scenes.append(make_scene(710, n_vehicles=2, n_roads=2, n_chunks=3, lane_ids=[0]))  # no interesting pair
    paths = write_dataset(str(tmp_path), scenes)
    rcfg = ref_harness.build_cfg(paths, 64, len(scenes))
    stub_policy = types.SimpleNamespace(real_time_rewards=False, model_path="x", cfg=rcfg, name="ctrl_sim",
                                        tilt_dict={"tilt": False})
    ref = PolicyEvaluator(rcfg, stub_policy)
    n_pairs = []
    for mode in ("one_agent", "two_agent"):
        random.seed(rcfg.eval.seed)
        want = []
        for f in range(len(scenes)):
            gt = get_ground_truth_states(rcfg, rcfg.nocturne_waymo_val_folder, ref.test_filenames, f, ref.dt, ref.steps)
            sim, scenario, vehicles = ref.load_scenario(rcfg.nocturne_waymo_val_folder, f)
            ref.vehicles_to_evaluate = get_moving_vehicles(scenario)
            if mode == "one_agent":
                v = ref.find_interesting_agent(vehicles, gt.copy())
                want.append(None if v is None else [int(v)])
            else:
                pr = ref.find_interesting_pair(vehicles, gt.copy())
                want.append(None if pr is None else [int(x) for x in pr])
            sim.reset()
        c = cfg.copy()
        c.eval = cfg.eval.copy()
        c.eval.eval_mode = mode
        ev = B200PolicyEvaluator(c, types.SimpleNamespace(model=types.SimpleNamespace(device="cpu")), scenes=scenes)
        _, ids, _, evs, _ = ev.select_scenes()
        assert ids == [f for f, w in enumerate(want) if w is not None]  # scenes without a pair are skipped
        assert evs == [w for w in want if w is not None], (mode, evs, want)
        n_pairs = [len(interesting_pairs(parse_scenario(s["json"]), [i for i, m in enumerate(parse_scenario(s["json"])["moving"]) if m]))
                   for s in scenes]
    assert sum(n > 2 for n in n_pairs) >= 3 and n_pairs[-1] == 0, n_pairs  # the draw is a real choice, and one scene has none
    c = cfg.copy()
    c.eval = cfg.eval.copy()
    c.eval.eval_mode = "three_agent"
    with pytest.raises(ValueError):
        B200PolicyEvaluator(c, types.SimpleNamespace(model=types.SimpleNamespace(device="cpu")), scenes=scenes).select_scenes()


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref")) or not os.path.isdir("/root/reference"),
                    reason="needs the reference tree and its built simulator (build container only)")
def test_reference_signature_adapter_runs_inside_the_stock_policy_evaluator(tmp_path, cfg):
    """``B200AutoregressivePolicy`` has the reference's constructor and reset / update_state / predict / act
    signatures, so the UNMODIFIED ``PolicyEvaluator.evaluate_policy()`` (its own Nocturne simulator and
    vehicle_data_dict) drives it.  Here the device is replaced by a recording stand-in (no GPU in this container; the GPU
    half is tests/test_gpu_parity.py::test_reference_signature_adapter_matches_the_batched_path): checks the call
    shapes, that the one-scene batch built from the reference's dicts equals the one our own loader builds from the
    scene JSON (ground truth, goals, sizes, evaluated set, focal order), the per-step state hand-over, and that what
    predict() writes through key_dict is what the evaluator then applies."""
    from ctrlsim_b200.batch import SceneBatch
    from ctrlsim_b200.policy_adapter import B200AutoregressivePolicy
    from ctrlsim_b200.scenario import parse_scenario
    from ctrlsim_b200.synth import make_scene, write_dataset
    from oracle import ref_harness, ref_shims
    ref_shims.install()
    from evaluators import PolicyEvaluator
    scenes = [make_scene(11, n_vehicles=6, n_roads=1, n_chunks=4), make_scene(12, n_vehicles=9, n_roads=2, n_chunks=3, frac_short=0.4)]
    paths = write_dataset(str(tmp_path), scenes)
    rcfg = ref_harness.build_cfg(paths, 4, len(scenes))  # at most 4 evaluated vehicles: random.sample is exercised
    steps = rcfg.nocturne.steps

    class Backend:
        def __init__(self):
            self.batches, self.calls = [], []

        def make_batch(self, cfg_, scene, scene_id, parsed, evaluated):
            b = SceneBatch(cfg_, [scene], [scene_id], "cpu", eval_threshold=len(evaluated), parsed=[parsed], evaluated_sets=[evaluated])
            self.batches.append((b, list(evaluated)))  # in the order random.sample drew them: it decides ties of the focal order
            return b

        def step(self, batch, t, states_t, actions_prev):
            n = batch.N
            assert states_t.shape == (n, 8) and np.isfinite(states_t).all()
            assert (actions_prev is None) == (t == 0)
            self.calls.append((len(self.batches) - 1, t, states_t.copy(), None if actions_prev is None else actions_prev.copy()))
            ev = self.batches[-1][1]
            nxt, rtg, act = np.zeros((n, 2)), -np.ones((n, 3), np.int64), -np.ones(n, np.int64)
            for v in ev:
                if states_t[v, 7]:
                    nxt[v] = (0.25 * (v + 1), 0.01 * (t % 7))
                    rtg[v], act[v] = (t, 35, 36), 524
            return nxt, rtg, act

    be = Backend()
    kd = {"next_acceleration": "next_acceleration", "next_steering": "next_steering", "rtgs": "rtgs"}
    td = {"tilt": True, "goal_tilt": 0, "veh_veh_tilt": 0, "veh_edge_tilt": 0}
    policy = B200AutoregressivePolicy(rcfg, "synthetic", None, True, True, True, False, False, False, False, kd, td, "ctrl_sim",
                                      1.0, False, 0.8, backend=be)
    recs = []
    ev = PolicyEvaluator(rcfg, policy)
    orig = ev.update_running_statistics
    ev.update_running_statistics = lambda d: (recs.append({k: {f: list(v[f]) if isinstance(v[f], list) else v[f] for f in ("acceleration", "steering", "rtgs", "existence")} for k, v in d.items()}), orig(d))[1]
    metrics, lines = ev.evaluate_policy()
    assert set(metrics) == {"goal", "collision_rate", "offroad_rate", "fde", "ade", "lin_speed_jsd", "ang_speed_jsd", "accel_jsd", "nearest_dist_jsd"}
    assert len(be.batches) == len(scenes) and len(be.calls) == len(scenes) * steps and len(recs) == len(scenes)
    assert [c[1] for c in be.calls] == list(range(steps)) * len(scenes)
    for k, sc in enumerate(scenes):
        b, evaluated = be.batches[k]
        p = parse_scenario(sc["json"], steps)
        assert b.t["scene_id"].tolist() == [k] and b.N == p["n"] and len(evaluated) == min(4, int(p["moving"].sum()))
        assert np.array_equal(b.t["gt"][0].numpy(), p["gt"].astype(np.float64)) and np.array_equal(b.t["gt_valid"][0].numpy(), p["gt_valid"])
        assert np.array_equal(b.t["goal"][0].numpy(), p["goal"])
        assert np.array_equal(b.t["veh_len"][0].numpy(), p["size"][:, 0]) and np.array_equal(b.t["veh_wid"][0].numpy(), p["size"][:, 1])
        ref_batch = SceneBatch(cfg, [sc], [k], "cpu", parsed=[p], evaluated_sets=[evaluated])
        for f in ("evaluated", "eval_order", "road_xy", "road_valid", "road_type", "n_poly"):
            assert np.array_equal(b.t[f].numpy(), ref_batch.t[f].numpy()), f
        # what predict() wrote through key_dict is what the evaluator applied (from the first controlled step, t = 9, on)
        d = recs[k]
        for v in evaluated:
            for t in range(rcfg.nocturne.history_steps - 1, steps):
                if d[v]["existence"][t]:
                    assert d[v]["acceleration"][t] == 0.25 * (v + 1) and d[v]["steering"][t] == 0.01 * (t % 7), (k, v, t)
                    assert np.allclose(d[v]["rtgs"][t], [t / 349 * 10, 35 / 349 * 100 - 10, 36 / 349 * 100 - 10])
        others = [v for v in range(p["n"]) if v not in evaluated]
        assert all((np.asarray(d[v]["rtgs"][t]) == 0).all() for v in others for t in range(steps))
        # the applied controls of step t-1 come back as the action history of step t
        for (kk, t, st, ap) in be.calls:
            if kk == k and t > 0:
                assert np.array_equal(ap[:, 0], [d[v]["acceleration"][t - 1] for v in range(p["n"])])


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref")) or not os.path.isdir("/root/reference"),
                    reason="needs the reference tree and its built simulator (build container only)")
def test_reference_signature_adapter_real_time_rewards_inside_the_stock_policy_evaluator(tmp_path):
    """SURVEY 8(b) x 8(f) N1: the adapter built with the cfgs/policy/dt.yaml switches inside the UNMODIFIED
    PolicyEvaluator.  The evaluator computes the dense reward and tracks the RTGs; the adapter must hand the device
    exactly vehicle_data_dict['rtgs'][t] at step t (policies/policy.py:91-93) and must not append to the series itself
    (autoregressive_policy.py:243-248 only runs with predict_rtgs).  Device replaced by a recording stand-in."""
    from ctrlsim_b200.batch import SceneBatch
    from ctrlsim_b200.policy_adapter import B200AutoregressivePolicy
    from ctrlsim_b200.synth import make_scene, write_dataset
    from oracle import ref_harness, ref_shims
    ref_shims.install()
    from evaluators import PolicyEvaluator
    scenes = [make_scene(11, n_vehicles=6, n_roads=1, n_chunks=4)]
    paths = write_dataset(str(tmp_path), scenes)
    rcfg = ref_harness.build_cfg(paths, 4, len(scenes))
    rcfg.model.decision_transformer, rcfg.model.predict_rtg, rcfg.model.predict_future_states = True, False, False
    steps = rcfg.nocturne.steps

    class Backend:
        def __init__(self):
            self.calls, self.evaluated = [], None

        def make_batch(self, cfg_, scene, scene_id, parsed, evaluated):
            self.evaluated = list(evaluated)
            return SceneBatch(cfg_, [scene], [scene_id], "cpu", eval_threshold=len(evaluated), parsed=[parsed], evaluated_sets=[evaluated])

        def step(self, batch, t, states_t, actions_prev, rtgs_t=None):
            assert rtgs_t is not None and rtgs_t.shape == (batch.N, 3)
            self.calls.append((t, rtgs_t.copy()))
            n = batch.N
            nxt, rtg, act = np.zeros((n, 2)), -np.ones((n, 3), np.int64), -np.ones(n, np.int64)
            for v in self.evaluated:
                if states_t[v, 7]:
                    nxt[v], act[v] = (0.5, 0.02), 524
            return nxt, rtg, act

    be = Backend()
    kd = {"next_acceleration": "next_acceleration", "next_steering": "next_steering", "rtgs": "rtgs"}
    td = {"tilt": False, "goal_tilt": 0, "veh_veh_tilt": 0, "veh_edge_tilt": 0}
    policy = B200AutoregressivePolicy(rcfg, "synthetic", None, True, False, False, True, False, True, False, kd, td, "dt",
                                      1.0, False, 0.8, backend=be)
    recs = []
    ev = PolicyEvaluator(rcfg, policy)
    orig = ev.update_running_statistics
    ev.update_running_statistics = lambda d: (recs.append({k: {"rtgs": [np.array(r) for r in v["rtgs"]], "dense": [np.array(r) for r in v["dense_reward"]],
                                                               "acceleration": list(v["acceleration"]), "existence": list(v["existence"])}
                                                           for k, v in d.items()}), orig(d))[1]
    metrics, _ = ev.evaluate_policy()
    assert "goal" in metrics and len(recs) == 1 and [c[0] for c in be.calls] == list(range(steps))
    d = recs[0]
    ids = list(d.keys())
    for t, rt in be.calls:
        want = np.array([d[v]["rtgs"][t] for v in ids])
        assert np.array_equal(rt, want), t
    for v in ids:
        assert len(d[v]["rtgs"]) == steps + 1  # one per update_vehicle_data_dict call; predict() appended nothing
        assert np.array_equal(d[v]["rtgs"][0], [10, 90, 90])
        for t in range(1, steps):
            assert np.array_equal(d[v]["rtgs"][t], d[v]["rtgs"][t - 1] - d[v]["dense"][t - 1])
    for v in be.evaluated:
        for t in range(rcfg.nocturne.history_steps - 1, steps):
            if d[v]["existence"][t]:
                assert d[v]["acceleration"][t] == 0.5


@pytest.mark.parametrize("name", ["plumbing", "sparse"])
def test_partition_scene_results_match_the_metrics_port(tmp_path, cfg, name):
    """N4: the per-partition ``scene_results`` JSON (policy_evaluator.py:578-593) - B200PolicyEvaluator.scene_results /
    write_partition_metrics restate update_running_statistics (:162-248) on the host from the device trace.  Fed with a
    reference fixture laid out as a trace, the lists equal the oracle's MetricsPort accumulators entry for entry."""
    import json
    import types
    import torch
    from ctrlsim_b200.evaluator import B200PolicyEvaluator
    from oracle.policy_port import MetricsPort
    g, spec, _ = load_golden(name)
    n, ev_ids = g["pos"].shape[0], [int(v) for v in g["evaluated"]]
    tr = {"tr_pos": g["pos"][None].astype(np.float32), "tr_vel": g["vel"][None].astype(np.float32),
          "tr_heading": g["heading"][None].astype(np.float32), "tr_exist": g["existence"][None].astype(np.uint8),
          "tr_action": np.stack([g["accel"], g["steer"]], -1)[None], "tr_reward": g["reward"][None].astype(np.float32),
          "tr_nearest": np.stack([g["nearest_dist"], g["gt_nearest_dist"]], -1)[None]}
    gt = np.concatenate([g["gt_pos"], g["gt_heading"][..., None], g["gt_speed"][..., None]], -1)[None]
    batch = types.SimpleNamespace(S=1, evaluated_ids=[ev_ids], trace=lambda: tr, t={"gt": torch.from_numpy(gt)})
    stub = types.SimpleNamespace(model=types.SimpleNamespace(device="cpu"), model_path=str(tmp_path / "model.ckpt"))
    ev = B200PolicyEvaluator(cfg, stub, scenes=[])
    res = ev.scene_results(batch)
    mp_ = MetricsPort(cfg)
    rec = {k: g[k] for k in ("existence", "reward", "pos", "gt_pos", "vel", "gt_speed", "heading", "gt_heading", "gt_accel", "accel",
                             "gt_nearest_dist", "nearest_dist")}
    mp_.add_scene(rec, ev_ids)
    assert res["goal_success"] == mp_.goal and res["collision"] == [float(x) for x in mp_.coll] and res["off_road"] == [float(x) for x in mp_.off]
    assert np.allclose(res["ade"], mp_.ade, rtol=0, atol=1e-6) and np.allclose(res["fde"], mp_.fde, rtol=0, atol=1e-6)  # fp32 trace positions
    for a, b in (("lin_speed_sim", "lin_sim"), ("lin_speed_gt", "lin_gt"), ("ang_speed_sim", "ang_sim"), ("ang_speed_gt", "ang_gt"),
                 ("accel_sim", "acc_sim"), ("accel_gt", "acc_gt"), ("nearest_dist_sim", "nd_sim"), ("nearest_dist_gt", "nd_gt")):
        assert len(res[a]) == len(mp_.samples[b])
        for x, y in zip(res[a], mp_.samples[b]):
            assert x.shape == y.shape and np.allclose(x, y, rtol=0, atol=1e-5), (a, np.abs(x - y).max())
    path = ev.write_partition_metrics(batch=batch)
    assert path.endswith(os.path.join("scene_results", "partition_0.json"))
    saved = json.load(open(path))
    assert set(saved) == {"goal_success", "ade", "fde", "accel_gt", "accel_sim", "ang_speed_gt", "ang_speed_sim", "lin_speed_gt",
                          "lin_speed_sim", "nearest_dist_gt", "nearest_dist_sim", "collision", "off_road"}  # the reference's keys
    assert len(saved["lin_speed_sim"]) == len(res["ade"]) and saved["goal_success"] == res["goal_success"]


def test_unsupported_simulator_switches_fail_loudly(cfg):
    """`collision_fix` and the rew_cfg switches are compiled into the kernels at their reference defaults
    (cfgs/config.yaml); any other value must raise instead of silently evaluating the default behaviour."""
    from ctrlsim_b200 import lib as L
    L.check_fixed_switches(cfg)
    for path, key, val in (("nocturne", "collision_fix", False), ("rew", "position_target", False),
                           ("rew", "shaped_goal_distance", False), ("rew", "heading_target", False)):
        c = cfg.copy()
        c.nocturne = cfg.nocturne.copy()
        if path == "nocturne":
            c.nocturne[key] = val
        else:
            c.nocturne["rew_cfg"] = dict(cfg.nocturne["rew_cfg"])
            c.nocturne["rew_cfg"][key] = val
        with pytest.raises(NotImplementedError):
            L.check_fixed_switches(c)


@needs_reference
def test_restated_config_equals_the_reference_yaml_tree(cfg):
    """ctrlsim_b200.config restates the constants of the reference's Hydra tree; load_reference_config composes the tree
    itself (cfgs/config.yaml + groups, without Hydra).  Every key default_config() defines must exist in the composed
    tree with the same value, except paths and three deliberate defaults (eval_mode, the two verbose flags); the same for
    dt_config(as_shipped=True) against `model=dt policy@eval.policy=dt` - which pins the `use_rtgs` typo of
    cfgs/policy/dt.yaml:11: as shipped, the DT baseline runs with use_rtg=False."""
    from ctrlsim_b200.config import dt_config, load_reference_config

    def flat(x, pre=""):
        out = {}
        for k, v in x.items():
            if isinstance(v, dict):
                out.update(flat(v, pre + k + "."))
            else:
                out[pre + k] = v
        return out

    def compare(mine, ref, allowed):
        fm, fr = flat(mine), flat(ref)
        assert [k for k in fm if k not in fr] == []
        diff = {k for k in fm if fm[k] != fr[k] and not (isinstance(fr[k], str) and not isinstance(fm[k], str) and float(fr[k]) == fm[k])}
        paths = {k for k in diff if isinstance(fr[k], str) and ("/" in fr[k])}
        assert diff - paths == allowed, (diff - paths, allowed)

    ref = load_reference_config("/root/reference/cfgs")
    compare(cfg, ref, {"eval.eval_mode", "eval.verbose", "eval_planner_adversary.verbose"})
    assert ref.eval.eval_mode == "one_agent" and ref.eval.policy.model == "ctrl_sim" and ref.nocturne.steps == 90
    assert ref.eval_planner_adversary.planner.goal_tilt == 10 and ref.eval_planner_adversary.adversary.veh_veh_tilt == -10
    assert isinstance(ref.nocturne["scenario"], dict) and not hasattr(ref.nocturne["scenario"], "copy_") and "hydra" not in ref
    ref_dt = load_reference_config("/root/reference/cfgs", groups={"model": "dt", "eval.policy": "dt"},
                                   overrides={"eval.eval_mode": "multi_agent"})
    assert ref_dt.eval.policy.use_rtg is False and ref_dt.eval.policy.use_rtgs is True  # the typo, as shipped
    assert ref_dt.model.decision_transformer is True and ref_dt.eval.policy.real_time_rewards is True
    compare(dt_config(as_shipped=True), ref_dt, {"eval.verbose", "eval_planner_adversary.verbose"})
    assert dt_config().eval.policy.use_rtg is True  # the intended mode stays the default of dt_config()


def test_rollout_port_dt_as_shipped_matches_reference_prefix():
    """cfgs/policy/dt.yaml AS SHIPPED (use_rtg=False through the `use_rtgs` typo): the unmodified reference evaluator feeds
    the DT network RTG (0, 0, 0) at every step while it keeps tracking the returns (tests/golden/rollout_dt_as_shipped.npz,
    oracle/make_golden.py).  Oracle port with use_rtg=False: same sampled actions, trajectories, dense reward, and the
    evaluator-side RTG series."""
    from ctrlsim_b200.config import dt_config
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    from oracle.model_port import ModelPort
    from oracle.policy_port import RolloutPort
    g, spec, _ = load_golden("dt_as_shipped")
    cfg = dt_config(as_shipped=True)
    fed = np.unique(g["in_9_rtgs_pass1"].reshape(-1, 3), axis=0)  # normalised zeros; (0, 0, 0) = padding rows
    assert len(fed) == 2 and np.allclose(fed, [[0.0, 0.0, 0.0], [0.0, 0.1, 0.1]])
    steps = 6
    port = RolloutPort(cfg, ModelPort(cfg, make_weights(cfg, **spec["weights"])), seed=0, eval_threshold=64,
                       predict_rtgs=False, discretize_rtgs=False, real_time_rewards=True, max_return=True, use_rtg=False)
    sc = make_scene(**spec["scene"])
    rec = port.run_scene(0, sc["json"], sc["preproc"], max_steps=steps)
    assert (rec["act_idx"][:steps] == g["act_idx"][:steps]).all()
    for k in ("pos", "vel", "heading", "existence", "dense_reward", "rtgs"):
        assert np.abs(rec[k][:, :steps] - g[k][:, :steps]).max() < 1e-9, k


def test_pipelined_evaluation_plumbing_on_cpu(cfg):
    """evaluate_policy(sub_batch_scenes=k): the producer thread hands over sub-batches in scene order whose scenes,
    global ids and evaluated-vehicle draws are exactly those of the one-batch selection (the seeded generator is
    walked once, in order), the per-sub-batch summaries are added, and an error in the producer reaches the caller."""
    import types
    from ctrlsim_b200.evaluator import B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    scenes = [make_scene(400 + i, n_vehicles=5 + 2 * i, n_roads=2, n_chunks=3) for i in range(7)]
    ids = [400 + i for i in range(7)]
    fake = types.SimpleNamespace(model=types.SimpleNamespace(device="cpu"))
    c = cfg.copy()
    c.eval = cfg.eval.copy()
    c.eval.multi_agent_eval_threshold = 4  # random.sample is exercised

    class Ev(B200PolicyEvaluator):
        def rollout(self, batch=None, max_steps=None):
            self.seen.append(batch)
            return batch

        def summarize(self, batch=None, local_only=False):
            assert local_only
            v = np.zeros(8 + 8 * 200)
            v[1] = batch.n_evaluated()
            v[4] = batch.S
            return v

        @staticmethod
        def metrics_from_summary(s, accel_bins=20):
            return {"n_agents": float(s[1]), "n_scenes": float(s[4])}

    one = B200PolicyEvaluator(c, fake, scenes=scenes, scene_ids=ids)
    _, want_ids, _, want_ev, _ = one.select_scenes()
    ev = Ev(c, fake, scenes=scenes, scene_ids=ids)
    ev.seen = []
    m, lines = ev.evaluate_policy(sub_batch_scenes=3)
    assert [b.S for b in ev.seen] == [3, 3, 1] and m == {"n_agents": float(sum(len(e) for e in want_ev)), "n_scenes": 7.0}
    got_ids = [int(i) for b in ev.seen for i in b.t["scene_id"].tolist()]
    got_ev = [e for b in ev.seen for e in b.evaluated_ids]
    assert got_ids == want_ids and got_ev == [sorted(e) for e in want_ev]
    assert ev.n_evaluated == sum(len(e) for e in want_ev) and len(lines) == 2
    bad = Ev(c, fake, scenes=scenes[:2] + [{"json": {"objects": "broken"}, "preproc": {}}], scene_ids=[1, 2, 3])
    bad.seen = []
    with pytest.raises(Exception):
        bad.evaluate_policy(sub_batch_scenes=2)


def test_pipelined_evaluation_releases_its_producer_when_the_consumer_fails(cfg):
    """An error while a sub-batch is rolled out must not leave the producer thread blocked on the hand-over queue."""
    import threading
    import types
    from ctrlsim_b200.evaluator import B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    scenes = [make_scene(420 + i, n_vehicles=4, n_roads=1, n_chunks=2) for i in range(8)]
    fake = types.SimpleNamespace(model=types.SimpleNamespace(device="cpu"))

    class Ev(B200PolicyEvaluator):
        def rollout(self, batch=None, max_steps=None):
            raise RuntimeError("device fault")

    before = threading.active_count()
    with pytest.raises(RuntimeError, match="device fault"):
        Ev(cfg, fake, scenes=scenes).evaluate_policy(sub_batch_scenes=1)
    assert threading.active_count() == before


def test_parser_paths_are_bit_identical():
    """parse_scenario's all-vehicles-at-once path (tracks of one length) and its per-vehicle loop give the same arrays,
    bit for bit and dtype for dtype - config-2 scenes, short / parked vehicles, replay scenes, the fixture scenes, and a
    scene without any vehicle valid at t = 0; a scene with tracks of different lengths takes the loop by itself."""
    import copy
    from ctrlsim_b200.scenario import parse_scenario
    from ctrlsim_b200.synth import make_replay_scene, make_scene
    scenes = [make_scene(i)["json"] for i in range(2)]
    scenes += [make_scene(100 + i, n_vehicles=3 + 5 * i, n_roads=1 + i % 3, n_chunks=2 + i % 4, frac_short=0.3, frac_parked=0.2)["json"] for i in range(6)]
    scenes += [make_replay_scene(i)["json"] for i in range(6)]
    scenes += [make_scene(**load_golden(name)[1]["scene"])["json"] for name in ("plumbing", "crowded", "sparse")]
    empty = copy.deepcopy(scenes[3])
    for o in empty["objects"]:
        o["valid"][0] = False
    scenes.append(empty)

    def same(a, b):
        assert a.keys() == b.keys()
        for k in a:
            if isinstance(a[k], list):
                assert len(a[k]) == len(b[k]) and all(np.array_equal(x, y) and x.dtype == y.dtype for x, y in zip(a[k], b[k])), k
            elif isinstance(a[k], np.ndarray):
                assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k
            else:
                assert a[k] == b[k], k

    for js in scenes:
        for steps in (90, 44):
            same(parse_scenario(js, steps), parse_scenario(js, steps, vectorize=False))
    ragged = copy.deepcopy(scenes[4])
    o = ragged["objects"][0]
    for k in ("position", "velocity", "heading", "valid"):
        o[k] = o[k] + [o[k][-1]] * 3
    same(parse_scenario(ragged, 90), parse_scenario(ragged, 90, vectorize=False))
    assert parse_scenario(ragged, 90)["n"] == parse_scenario(scenes[4], 90)["n"]
