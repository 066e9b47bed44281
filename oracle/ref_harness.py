"""ORACLE (test infrastructure only): run the UNMODIFIED reference evaluator on synthetic scenes and record it.

``run_reference(scenes, ...)`` writes the scenes in the reference's on-disk formats, builds the reference ``CtRLSim``
model with our deterministic weights, wraps it in the reference ``AutoregressivePolicy`` + ``PolicyEvaluator`` and
calls ``evaluate_policy()`` (evaluators/policy_evaluator.py:426-595) on the real ``nocturne_cpp`` built by
oracle/Makefile.  The only behavioural change is the sampler: ``torch.multinomial`` is replaced by the explicit
sampler of oracle/sampler.py (see its header for why).  Everything observable is recorded per scene:
trajectories, applied actions, rewards, sampled indices, focal groups, (optionally) logits, and the final metrics.
"""
from __future__ import annotations

import os
import tempfile
import types

import numpy as np
import torch

from ctrlsim_b200.config import default_config
from ctrlsim_b200.synth import write_dataset
from ctrlsim_b200.weights import make_weights

from . import ref_shims, sampler


class _Ctx:
    seed = 0
    scene = 0
    step = 0
    agent = 0
    comp = 0
    action_queue = []
    logits_of = {}
    rec = None


def _softmax_shim(x, dim=0, **kw):
    out = torch.softmax(x, dim=dim)
    _Ctx.logits_of[id(out)] = (out, x.detach().clone())
    return out


def _multinomial_shim(dist, n, *a, **k):
    assert n == 1
    ent = _Ctx.logits_of.pop(id(dist), None)
    assert ent is not None and ent[0] is dist, "multinomial called on a tensor that did not come from F.softmax"
    x = ent[1].to(torch.float32).numpy()  # RTG: float64 (logit + tilt) -> fp32; action: already fp32
    if x.shape[0] == 1000:
        agent = _Ctx.action_queue.pop(0)
        comp = sampler.COMP_ACTION
    else:
        agent, comp = _Ctx.agent, _Ctx.comp
        _Ctx.comp += 1
    idx = sampler.sample_from_x(x, _Ctx.seed, _Ctx.scene, agent, _Ctx.step, comp)
    rec = _Ctx.rec
    if comp == sampler.COMP_ACTION:
        rec["act_idx"][_Ctx.step, agent] = idx
    else:
        rec["rtg_idx"][_Ctx.step, agent, comp] = idx
    return torch.tensor([idx], dtype=torch.int64)


def build_cfg(paths, eval_threshold=64, num_files=1000):
    cfg = default_config()
    cfg.dataset_root = paths["dataset_root"]
    cfg.nocturne_waymo_val_folder = paths["nocturne_waymo_val_folder"]
    cfg.dataset.waymo.preprocess_dir = paths["preprocess_dir"]
    cfg.eval.eval_mode = "multi_agent"
    cfg.eval.multi_agent_eval_threshold = eval_threshold
    cfg.eval.num_files_to_evaluate = num_files
    cfg.eval.verbose = False
    return cfg


def build_reference_model(cfg, weights):
    ref_shims.install()
    from models import CtRLSim
    model = CtRLSim(cfg)
    sd = {k: torch.from_numpy(v.copy()) for k, v in weights.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("causal_mask" in m for m in missing), missing
    return model.eval()


def run_reference(scenes, weights=None, seed=0, tilts=(0, 0, 0), temperature=1.0, eval_threshold=64,
                  steps=90, logit_steps=(), workdir=None, cfg_hook=None, policy_opts=None):
    """Returns (metrics_dict, [per-scene record dict]).

    ``policy_opts`` overrides the AutoregressivePolicy constructor arguments (cfgs/policy/*.yaml), e.g. the DT baseline
    of cfgs/policy/dt.yaml: dict(predict_rtgs=False, discretize_rtgs=False, real_time_rewards=True, max_return=True,
    name="dt", tilt_dict={"tilt": False, ...}) together with a ``cfg_hook`` that selects cfgs/model/dt.yaml.  With
    real_time_rewards the records also hold the dense reward series (evaluators/evaluator.py:106-140)."""
    ref_shims.install()
    import policies.policy as ref_policy_mod
    import policies.autoregressive_policy as ref_ar_mod
    from policies import AutoregressivePolicy
    from evaluators import PolicyEvaluator

    tmp = workdir or tempfile.mkdtemp(prefix="ctrlsim_ref_")
    paths = write_dataset(tmp, scenes)
    cfg = build_cfg(paths, eval_threshold, len(scenes))
    cfg.nocturne.steps = steps
    if cfg_hook:
        cfg_hook(cfg)
    if weights is None:
        weights = make_weights(cfg)
    model = build_reference_model(cfg, weights)

    # --- explicit sampler -------------------------------------------------------------------------------------
    Fshim = types.SimpleNamespace(softmax=_softmax_shim)
    ref_policy_mod.F = Fshim
    ref_ar_mod.F = Fshim
    orig_multinomial = torch.multinomial
    torch.multinomial = _multinomial_shim
    _Ctx.seed = seed

    pkw = dict(
        cfg=cfg, model_path="synthetic", model=model, use_rtg=True, predict_rtgs=True, discretize_rtgs=True,
        real_time_rewards=False, privileged_return=False, max_return=False, min_return=False,
        key_dict={"next_acceleration": "next_acceleration", "next_steering": "next_steering", "rtgs": "rtgs"},
        tilt_dict={"tilt": True, "goal_tilt": tilts[0], "veh_veh_tilt": tilts[1], "veh_edge_tilt": tilts[2]},
        name="ctrl_sim", action_temperature=temperature, nucleus_sampling=False, nucleus_threshold=0.8)
    pkw.update(policy_opts or {})
    policy = AutoregressivePolicy(**pkw)
    passes = 2 if pkw["predict_rtgs"] else 1  # model calls per focal group (autoregressive_policy.py:190,210)
    evaluator = PolicyEvaluator(cfg, policy)

    records = []
    n_max = max(len(s["json"]["objects"]) for s in scenes)

    def new_record(scene_idx):
        n = len(scenes[scene_idx]["json"]["objects"])
        return {"scene": scene_idx, "n": n,
                "rtg_idx": -np.ones((steps, n, 3), np.int32), "act_idx": -np.ones((steps, n), np.int32),
                "groups": [[] for _ in range(steps)], "logits": {}}

    # scene id: load_scenario(file_path, file) is called once per scene before the step loop
    orig_load = evaluator.load_scenario

    def load_scenario(file_path, file):
        _Ctx.scene = int(file)
        _Ctx.rec = new_record(int(file))
        return orig_load(file_path, file)

    evaluator.load_scenario = load_scenario

    orig_get_data = policy.get_data

    def get_data(gt_data_dict, preproc_data, dset, vehicles_to_evaluate, t):
        out = orig_get_data(gt_data_dict, preproc_data, dset, vehicles_to_evaluate, t)
        motion_datas, dead, new_idx, data_veh_ids = out
        _Ctx.step = int(t)
        _Ctx.action_queue = [int(v) for f in motion_datas.keys() for v in data_veh_ids[f]]
        for f in motion_datas.keys():
            members = -np.ones(24, np.int32)
            for old, new in new_idx[f].items():
                members[new] = int(old)
            _Ctx.rec["groups"][t].append({"focal": int(f), "members": members,
                                          "served": [int(v) for v in data_veh_ids[f]]})
        _Ctx.cur_group = 0
        return out

    policy.get_data = get_data

    orig_ppr = policy.process_predicted_rtg

    def process_predicted_rtg(rtg_logits, token_index, veh_id, *a, **k):
        _Ctx.agent, _Ctx.comp = int(veh_id), 0
        return orig_ppr(rtg_logits, token_index, veh_id, *a, **k)

    policy.process_predicted_rtg = process_predicted_rtg

    # logits (and the exact model inputs) at selected steps: wrap the model call
    orig_forward = model.forward
    call_no = {"n": 0, "step": -1}

    def forward(data, eval=False):
        preds = orig_forward(data, eval)
        t = _Ctx.step
        if t in logit_steps:
            if call_no["step"] != t:
                call_no["step"], call_no["n"] = t, 0
            g, p = divmod(call_no["n"], passes)
            call_no["n"] += 1
            ti = t if t < cfg.dataset.waymo.train_context_length else -1
            ent = _Ctx.rec["logits"].setdefault((t, g), {})
            if p == 0:
                if passes == 2:
                    ent["rtg_logits"] = preds["rtg_preds"][0, :, ti].detach().numpy().astype(np.float32).copy()
                ent["inputs"] = {
                    "agent_states": data["agent"].agent_states[0].numpy().copy(),
                    "agent_types": data["agent"].agent_types[0].numpy().copy(),
                    "goals": data["agent"].goals[0].numpy().copy(),
                    "actions": data["agent"].actions[0].numpy().copy(),
                    "rtgs_pass1": data["agent"].rtgs[0].numpy().copy(),
                    "timesteps": data["agent"].timesteps[0].numpy().copy(),
                    "road_points": data["map"].road_points[0].numpy().astype(np.float32).copy(),
                    "road_types": data["map"].road_types[0].numpy().astype(np.float32).copy(),
                }
            if p == passes - 1:
                ent["action_logits"] = preds["action_preds"][0, :, ti].detach().numpy().astype(np.float32).copy()
                ent["rtgs_pass2"] = data["agent"].rtgs[0].numpy().copy()
        return preds

    model.forward = forward

    orig_stats = evaluator.update_running_statistics

    def update_running_statistics(data_dict):
        rec = _Ctx.rec
        ids = sorted(data_dict.keys())
        T = steps + 1

        def arr(key, sub=None):
            if sub is None:
                return np.array([[data_dict[v][key][t] for t in range(T)] for v in ids], dtype=np.float64)
            return np.array([[[data_dict[v][key][t][s] for s in sub] for t in range(T)] for v in ids], np.float64)

        rec["veh_ids"] = np.array(ids, np.int32)
        rec["pos"] = arr("position", ("x", "y"))
        rec["vel"] = arr("velocity", ("x", "y"))
        rec["heading"] = arr("heading")
        rec["existence"] = arr("existence")
        rec["accel"] = arr("acceleration")
        rec["steer"] = arr("steering")
        rec["reward"] = np.array([[data_dict[v]["reward"][t] for t in range(T)] for v in ids], np.float64)
        rec["rtgs"] = np.array([[data_dict[v]["rtgs"][t] for t in range(steps)] for v in ids], np.float64)
        if policy.real_time_rewards:
            rec["dense_reward"] = np.array([[data_dict[v]["dense_reward"][t] for t in range(T)] for v in ids], np.float64)
        rec["nearest_dist"] = arr("nearest_dist")
        rec["gt_nearest_dist"] = arr("gt_nearest_dist")
        rec["gt_pos"] = arr("gt_position", ("x", "y"))
        rec["gt_heading"] = arr("gt_heading")
        rec["gt_speed"] = arr("gt_speed")
        rec["gt_accel"] = arr("gt_acceleration")
        rec["goal"] = np.array([[data_dict[v]["goal_position"]["x"], data_dict[v]["goal_position"]["y"],
                                 data_dict[v]["goal_heading"], data_dict[v]["goal_speed"]] for v in ids], np.float64)
        rec["size"] = np.array([[data_dict[v]["length"], data_dict[v]["width"]] for v in ids], np.float64)
        rec["evaluated"] = np.array(sorted(int(v) for v in evaluator.vehicles_to_evaluate), np.int32)
        records.append(rec)
        return orig_stats(data_dict)

    evaluator.update_running_statistics = update_running_statistics

    try:
        with torch.no_grad():
            metrics, _ = evaluator.evaluate_policy()
    finally:
        torch.multinomial = orig_multinomial
    return {k: float(v) for k, v in metrics.items()}, records


def run_reference_planner_adversary(scenes, pairs, adv_positions=None, weights=None, seeds=(1, 2),
                                    tilts_planner=(10, 10, 10), tilts_adversary=(0, -10, 0), cat=False, steps=90,
                                    workdir=None):
    """The UNMODIFIED ``PlannerAdversaryEvaluator.evaluate_planner_adversary()``
    (evaluators/planner_adversary_evaluator.py:466-593) on synthetic scenes.

    ``pairs``: per scene (ego, adversary) JSON object indices; ``adv_positions``: per scene the CAT trajectory [steps+1, 2]
    (the evaluator requires one in its table even when the adversary is a CtRL-Sim policy, :447-448).
    ``cat``: the adversary follows that trajectory instead of a policy (cfgs/policy/cat.yaml).
    Returns (metrics, [per-scene record]); a record holds the world arrays plus 'planner' / 'adversary' sub-records
    (sampled bins, RTGs, focal groups) like oracle/planner_adversary_port.py."""
    import pickle
    import shutil
    ref_shims.install()
    import policies.policy as ref_policy_mod
    import policies.autoregressive_policy as ref_ar_mod
    from policies import AutoregressivePolicy
    from evaluators import PlannerAdversaryEvaluator

    tmp = workdir or tempfile.mkdtemp(prefix="ctrlsim_ref_pa_")
    paths = write_dataset(tmp, scenes)
    shutil.copytree(os.path.join(paths["preprocess_dir"], "test"), os.path.join(paths["preprocess_dir"], "val_interactive"),
                    dirs_exist_ok=True)
    table = {}
    for k, sc in enumerate(scenes):
        ent = {"nocturne_path": "x" * 66 + sc["name"] + ".json", "nocturne_sdc_id": int(pairs[k][0]),
               "nocturne_adversary_id": int(pairs[k][1])}
        if adv_positions is not None and adv_positions[k] is not None:
            ent["adv_traj"] = np.asarray(adv_positions[k], np.float64)
        table[k] = ent
    dict_path = os.path.join(tmp, "eval_planner_dict.pkl")
    with open(dict_path, "wb") as f:
        pickle.dump(table, f)
    cfg = build_cfg(paths, 2, len(scenes))
    cfg.nocturne.steps = steps
    cfg.nocturne_waymo_val_interactive_folder = paths["nocturne_waymo_val_folder"]
    cfg.cat.dict_path = dict_path
    cfg.eval_planner_adversary.num_files_to_evaluate = len(scenes)
    cfg.eval_planner_adversary.verbose = False
    cfg.eval_planner_adversary.visualize = False
    if weights is None:
        weights = make_weights(cfg)
    model = build_reference_model(cfg, weights)

    Fshim = types.SimpleNamespace(softmax=_softmax_shim)
    ref_policy_mod.F = Fshim
    ref_ar_mod.F = Fshim
    orig_multinomial = torch.multinomial
    torch.multinomial = _multinomial_shim

    def make_policy(name, tilts, keys):
        return AutoregressivePolicy(
            cfg=cfg, model_path="synthetic", model=model, use_rtg=True, predict_rtgs=True, discretize_rtgs=True,
            real_time_rewards=False, privileged_return=False, max_return=False, min_return=False, key_dict=keys,
            tilt_dict={"tilt": True, "goal_tilt": tilts[0], "veh_veh_tilt": tilts[1], "veh_edge_tilt": tilts[2]},
            name=name, action_temperature=1.0, nucleus_sampling=False, nucleus_threshold=0.8)

    planner = make_policy("ctrl_sim", tilts_planner, {"next_acceleration": "next_planner_acceleration",
                                                      "next_steering": "next_planner_steering", "rtgs": "planner_rtgs"})
    if cat:
        adversary = types.SimpleNamespace(name="cat", real_time_rewards=False, model_path="cat", reset=lambda d: None)
    else:
        adversary = make_policy("ctrl_sim", tilts_adversary, {"next_acceleration": "next_adversary_acceleration",
                                                              "next_steering": "next_adversary_steering",
                                                              "rtgs": "adversary_rtgs"})
    evaluator = PlannerAdversaryEvaluator(cfg, planner, adversary)
    records = []
    scene = {"rec": None}

    def policy_record(n):
        return {"rtg_idx": -np.ones((steps, n, 3), np.int32), "act_idx": -np.ones((steps, n), np.int32),
                "groups": [[] for _ in range(steps)], "logits": {}}

    orig_load = evaluator.load_scenario

    def load_scenario(file_path, file):
        n = len(scenes[int(file)]["json"]["objects"])
        _Ctx.scene = int(file)
        scene["rec"] = {"scene": int(file), "n": n, "planner": policy_record(n),
                        "adversary": None if cat else policy_record(n)}
        return orig_load(file_path, file)

    evaluator.load_scenario = load_scenario

    def instrument(policy, role, seed):
        orig_get_data = policy.get_data

        def get_data(gt_data_dict, preproc_data, dset, vehicles_to_evaluate, t):
            out = orig_get_data(gt_data_dict, preproc_data, dset, vehicles_to_evaluate, t)
            motion_datas, dead, new_idx, data_veh_ids = out
            _Ctx.seed, _Ctx.step, _Ctx.rec = seed, int(t), scene["rec"][role]
            _Ctx.action_queue = [int(v) for f in motion_datas.keys() for v in data_veh_ids[f]]
            for f in motion_datas.keys():
                members = -np.ones(24, np.int32)
                for old, new in new_idx[f].items():
                    members[new] = int(old)
                _Ctx.rec["groups"][t].append({"focal": int(f), "members": members,
                                              "served": [int(v) for v in data_veh_ids[f]]})
            return out

        policy.get_data = get_data
        orig_ppr = policy.process_predicted_rtg

        def process_predicted_rtg(rtg_logits, token_index, veh_id, *a, **k):
            _Ctx.agent, _Ctx.comp = int(veh_id), 0
            return orig_ppr(rtg_logits, token_index, veh_id, *a, **k)

        policy.process_predicted_rtg = process_predicted_rtg

    instrument(planner, "planner", seeds[0])
    if not cat:
        instrument(adversary, "adversary", seeds[1])

    orig_stats = evaluator.update_running_statistics

    def update_running_statistics(data_dict):
        rec = scene["rec"]
        ids = sorted(data_dict.keys())
        T = steps + 1

        def arr(key, sub=None):
            if sub is None:
                return np.array([[data_dict[v][key][t] for t in range(T)] for v in ids], dtype=np.float64)
            return np.array([[[data_dict[v][key][t][s] for s in sub] for t in range(T)] for v in ids], np.float64)

        rec["veh_ids"] = np.array(ids, np.int32)
        rec["pos"], rec["vel"] = arr("position", ("x", "y")), arr("velocity", ("x", "y"))
        for dst, src in (("heading", "heading"), ("existence", "existence"), ("accel", "acceleration"),
                         ("steer", "steering"), ("nearest_dist", "nearest_dist"), ("gt_nearest_dist", "gt_nearest_dist"),
                         ("gt_heading", "gt_heading"), ("gt_speed", "gt_speed"), ("gt_accel", "gt_acceleration")):
            rec[dst] = arr(src)
        rec["gt_pos"] = arr("gt_position", ("x", "y"))
        rec["reward"] = np.array([[data_dict[v]["reward"][t] for t in range(T)] for v in ids], np.float64)
        rec["planner"]["rtgs"] = np.array([[data_dict[v]["planner_rtgs"][t] for t in range(steps)] for v in ids], np.float64)
        if not cat:
            rec["adversary"]["rtgs"] = np.array([[data_dict[v]["adversary_rtgs"][t] for t in range(steps)] for v in ids],
                                                np.float64)
        rec["size"] = np.array([[data_dict[v]["length"], data_dict[v]["width"]] for v in ids], np.float64)
        rec["ego"], rec["adv"] = int(evaluator.ego_vehicle), int(evaluator.adversary_vehicle)
        records.append(rec)
        return orig_stats(data_dict)

    evaluator.update_running_statistics = update_running_statistics
    try:
        with torch.no_grad(), np.errstate(invalid="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                metrics, _ = evaluator.evaluate_planner_adversary()
    finally:
        torch.multinomial = orig_multinomial
    return {k: float(v) for k, v in metrics.items()}, records
