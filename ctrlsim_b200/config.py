"""Configuration for the rollout hot path.

The reference drives everything from a Hydra/OmegaConf tree (cfgs/config.yaml + groups).  Hydra is not a dependency
of this package; the handful of constants the rollout path reads are restated here as a plain attribute-dict with the
same key paths (``cfg.dataset.waymo.*``, ``cfg.model.*``, ``cfg.nocturne.*``, ``cfg.eval.*``) so that code written
against the reference config keeps working.

Values follow (reference file:line):
  cfgs/dataset/waymo/base.yaml:4-43   context length, thresholds, action/RTG ranges and discretisation, caps
  cfgs/model/base.yaml:1-18, cfgs/model/ctrl_sim.yaml:4-9   model sizes
  cfgs/config.yaml:41-90              nocturne steps/dt/history, scenario dict, reward config
  cfgs/eval/base.yaml:5-15, cfgs/policy/ctrl_sim.yaml:5-13  eval + policy defaults
  cfgs/eval_planner_adversary/base.yaml:1-10, cfgs/policy/ctrl_sim_{planner,adversary}.yaml  planner-vs-adversary defaults
"""
from __future__ import annotations

import copy


class AttrDict(dict):
    """dict with attribute access whose ``copy()`` stays an AttrDict.

    The reference calls ``cfg.copy()`` (policies/policy.py:24), ``cfg.nocturne['rew_cfg']``
    (evaluators/policy_evaluator.py:151) and ``rew_cfg.get(...)`` (utils/sim.py:111), so both item and attribute
    access must work and the scenario sub-dict must be a real ``dict`` (cfgs/config.py:18-21).
    """

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        return AttrDict({k: (v.copy() if isinstance(v, AttrDict) else copy.copy(v)) for k, v in self.items()})

    @staticmethod
    def wrap(d):
        if isinstance(d, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in d.items()})
        return d


def default_config(wide: bool = False) -> AttrDict:
    """``wide``: SURVEY 8(d) config 2 "wide" variant - the caps of cfgs/dataset/waymo/base.yaml:38-39 raised to 64 agents
    and 256 polylines per focal group, so that one group covers a whole 64-vehicle / 256-polyline scene (6144 decoder
    tokens, 320 memory tokens); everything else as the reference default."""
    waymo = dict(
        train_context_length=32, num_agent_types=5, num_road_types=8, map_attr=2, k_attr=7,
        agent_dist_threshold=60.0, map_dist_threshold=100.0, max_timestep=90,
        parked_car_velocity_threshold=0.05,
        max_accel=10.0, min_accel=-10.0, max_steer=0.7, min_steer=-0.7,
        max_veh_veh_distance=15.0, dist_to_road_edge_scaling_factor=15.0,
        veh_veh_collision_rew_multiplier=10.0, veh_edge_collision_rew_multiplier=10.0,
        pos_goal_shaped_min=0, pos_goal_shaped_max=0.2, pos_target_achieved_rew_multiplier=10.0,
        moving_threshold=0.05,
        min_rtg_pos=0, max_rtg_pos=10, min_rtg_yaw=0, max_rtg_yaw=110, min_rtg_vel=0, max_rtg_vel=110,
        min_rtg_veh=-10, max_rtg_veh=90, min_rtg_road=-10, max_rtg_road=90,
        max_num_agents=24, max_num_road_polylines=200, max_num_road_pts_per_polyline=100,
        accel_discretization=20, steer_discretization=50, rtg_discretization=350,
        preprocess=True, preprocess_real_data=False, preprocess_simulated_data=False,
        goal_dim=5, remove_shaped_goal=True, remove_shaped_veh_reward=False, remove_shaped_edge_reward=False,
        dataset_path="", preprocess_dir="", simulated_dataset="", simulated_dataset_preprocessed_dir="",
    )
    model = dict(
        hidden_dim=256, map_attr=3, num_road_types=8, no_actions=False, num_heads=8, num_reward_components=3,
        dim_feedforward=1024, dropout=0.1, state_dim=12, use_map=True, goal_dropout=0.1, max_pool_map=True,
        supervise_moving=True, predict_rtg=True, attend_own_return_action=False, trajeglish=False, il=False,
        ctg_plus_plus=False, decision_transformer=False,
        num_transformer_encoder_layers=2, num_decoder_layers=4, predict_future_states=True,
        local_frame_predictions=False, loss_action_coef=1.0, encode_initial_state=True,
    )
    nocturne = dict(
        collision_fix=True, steps=90, dt=0.1, history_steps=10,
        scenario=dict(
            start_time=0, allow_non_vehicles=False, moving_threshold=0.2, speed_threshold=0.05,
            max_visible_objects=16, max_visible_road_points=1000, max_visible_traffic_lights=20,
            max_visible_stop_signs=4, sample_every_n=1, road_edge_first=False,
        ),
        rew_cfg=dict(
            shared_reward=False, goal_tolerance=0.5, reward_scaling=1.0, collision_penalty=0,
            shaped_goal_distance_scaling=0.2, shaped_goal_distance=True, goal_distance_penalty=False,
            position_target=True, position_target_tolerance=1.0, speed_target=True, speed_target_tolerance=1.0,
            heading_target=True, heading_target_tolerance=0.3,
        ),
    )
    policy = dict(
        run_name="ctrl_sim", model_path="", model="ctrl_sim",
        veh_veh_tilt=0, veh_edge_tilt=0, goal_tilt=0,
        action_temperature=1.0, nucleus_sampling=False, nucleus_threshold=0.8,
        use_rtg=True, predict_rtgs=True, discretize_rtgs=True, real_time_rewards=False,
        privileged_return=False, max_return=False, min_return=False,
    )
    evalc = dict(
        movie_path="", visualize=False, history_steps=10, interesting_traj_len_threshold=60,
        interesting_goal_dist_threshold=10, interesting_timestep_diff_threshold=20,
        multi_agent_eval_threshold=8, num_files_to_evaluate=1000, eval_mode="multi_agent", verbose=False,
        seed=0, partitions=1, policy=policy,
    )
    # cfgs/eval_planner_adversary/base.yaml, cfgs/policy/ctrl_sim_planner.yaml, cfgs/policy/ctrl_sim_adversary.yaml
    epa = dict(
        history_steps=10, verbose=False, seed=0, visualize=False, visualization_data_path="",
        num_files_to_evaluate=1000,
        planner=dict(policy, goal_tilt=10, veh_veh_tilt=10, veh_edge_tilt=10),
        adversary=dict(policy, goal_tilt=0, veh_veh_tilt=-10, veh_edge_tilt=0),
    )
    train = dict(seed=0, max_steps=200000, warmup_steps=500, lr=5e-4, weight_decay=1e-4, finetuning=False)
    cfg = AttrDict.wrap(dict(
        dataset_root="", project_root="", nocturne_waymo_val_folder="", nocturne_waymo_val_interactive_folder="",
        dataset=dict(waymo=waymo), model=model, nocturne=nocturne, eval=evalc, train=train,
        eval_planner_adversary=epa, cat=dict(dict_path=""),
    ))
    if wide:
        cfg.dataset.waymo.max_num_agents = 64
        cfg.dataset.waymo.max_num_road_polylines = 256
    # the scenario dict is handed to pybind as-is by the reference: keep it a plain dict
    cfg.nocturne["scenario"] = dict(nocturne["scenario"])
    cfg.nocturne["rew_cfg"] = AttrDict(nocturne["rew_cfg"])
    return cfg


def dt_config(wide: bool = False, as_shipped: bool = False) -> AttrDict:
    """The decision-transformer baseline: cfgs/model/dt.yaml on top of cfgs/model/ctrl_sim.yaml
    (continuous RTG inputs, (rtg, state, action) token order, no RTG head, no future-state head) and cfgs/policy/dt.yaml
    (RTGs are not predicted but tracked in real time: they start at the maximum return and are decremented by the dense
    reward of every step)."""
    cfg = default_config(wide)
    cfg.model.decision_transformer, cfg.model.predict_rtg, cfg.model.predict_future_states = True, False, False
    # NOTE: ``use_rtg=True`` is the evident intent of cfgs/policy/dt.yaml:11, which spells the key ``use_rtgs``; what Hydra
    # composes from the reference's files AS SHIPPED is use_rtg=False (cfgs/policy/base.yaml:1; load_reference_config
    # reproduces that), and then the network is fed RTG (0, 0, 0).  Both are implemented: dt_config(as_shipped=True).
    cfg.eval.policy.update(run_name="dt", model="dt", use_rtg=not as_shipped, predict_rtgs=False, discretize_rtgs=False,
                           real_time_rewards=True, max_return=True)
    for k in ("veh_veh_tilt", "veh_edge_tilt", "goal_tilt"):  # cfgs/policy/dt.yaml has no tilt keys
        cfg.eval.policy.pop(k, None)
    return cfg



# ---- reading the reference's own configuration tree --------------------------------------------------------------------
def _load_group_file(cfgs_dir: str, group: str, option: str) -> dict:
    """One file of a config group with its own ``defaults`` list composed first (Hydra semantics for the forms the
    reference uses: ``- base`` = sibling option of the same group merged underneath; ``- /grp@key: opt`` = option
    ``opt`` of the absolute group ``grp`` placed at ``key`` inside this file's package)."""
    import os
    import yaml
    with open(os.path.join(cfgs_dir, group, option + ".yaml")) as f:
        body = yaml.safe_load(f) or {}
    out: dict = {}
    for d in body.pop("defaults", None) or []:
        if isinstance(d, str):
            if d != "_self_":
                _deep_merge(out, _load_group_file(cfgs_dir, group, d))
        else:
            (k, opt), = d.items()
            if k.startswith("/"):
                grp, _, at = k[1:].partition("@")
                _deep_merge(out, {at or grp.replace("/", "."): _load_group_file(cfgs_dir, grp, opt)})
            else:
                sub, _, at = k.partition("@")
                _deep_merge(out, {at or sub: _load_group_file(cfgs_dir, os.path.join(group, sub), opt)})
    _deep_merge(out, body)
    return out


def _deep_merge(dst: dict, src: dict) -> dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _deep_merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


def _set_path(tree: dict, path: str, value):
    keys = path.split(".")
    for k in keys[:-1]:
        tree = tree.setdefault(k, {})
    if isinstance(value, dict) and isinstance(tree.get(keys[-1]), dict):
        _deep_merge(tree[keys[-1]], value)
    else:
        tree[keys[-1]] = value


def _resolve(tree: dict):
    """``${a.b.c}`` interpolations against the root (OmegaConf); resolver calls such as ``${now:...}`` stay as they are."""
    import re
    pat = re.compile(r"\$\{([A-Za-z0-9_.]+)\}")

    def get(path):
        node = tree
        for k in path.split("."):
            node = node[k]
        return node

    def walk(node):
        for k, v in (node.items() if isinstance(node, dict) else enumerate(node)):
            if isinstance(v, (dict, list)):
                walk(v)
            elif isinstance(v, str):
                for _ in range(8):
                    m = pat.fullmatch(v)
                    if m:  # a whole-value reference keeps the referenced type
                        v = get(m.group(1))
                        if not isinstance(v, str):
                            break
                        continue
                    new = pat.sub(lambda mm: str(get(mm.group(1))), v)
                    if new == v:
                        break
                    v = new
                node[k] = v
    walk(tree)


def load_reference_config(cfgs_dir: str, groups: dict | None = None, overrides: dict | None = None) -> AttrDict:
    """Compose the reference's Hydra tree (``<cfgs_dir>/config.yaml`` + its config groups) into the attribute-dict the
    rollout path reads - what ``@hydra.main(config_path=CONFIG_PATH, config_name="config")`` hands to ``eval_sim.py``
    (cfgs/config.yaml:8-14, eval_sim.py:7-8), without Hydra.

    ``groups``: group choices that replace the defaults list's, e.g. ``{"model": "dt", "eval.policy": "dt"}`` for
    ``python eval_sim.py model=dt policy@eval.policy=dt`` (keys: the package the group lands in; the group directory of
    ``eval.policy`` is ``policy``).  ``overrides``: dotted key -> value, applied last (``{"eval.eval_mode":
    "multi_agent", "eval.policy.goal_tilt": 10}``)."""
    import os
    import yaml
    with open(os.path.join(cfgs_dir, "config.yaml")) as f:
        root = yaml.safe_load(f)
    tree: dict = {}
    chosen = dict(groups or {})
    for d in root.pop("defaults", []):
        (grp, opt), = d.items()
        pkg = grp.replace("/", ".")
        _set_path(tree, pkg, _load_group_file(cfgs_dir, grp, chosen.pop(pkg, opt)))
    for pkg, opt in chosen.items():  # nested choices such as eval.policy / eval_planner_adversary.planner -> group "policy"
        grp = {"policy": "policy", "planner": "policy", "adversary": "policy"}.get(pkg.split(".")[-1], pkg.replace(".", "/"))
        node = tree
        for k in pkg.split(".")[:-1]:
            node = node.setdefault(k, {})
        node[pkg.split(".")[-1]] = _load_group_file(cfgs_dir, grp, opt)  # a group choice REPLACES the default option
    _deep_merge(tree, root)
    tree.pop("hydra", None)
    for k, v in (overrides or {}).items():
        _set_path(tree, k, v)
    _resolve(tree)
    cfg = AttrDict.wrap(tree)
    # the scenario dict is handed to pybind as-is by the reference: keep it a plain dict (cfgs/config.py:18-21)
    cfg.nocturne["scenario"] = dict(tree["nocturne"]["scenario"])
    return cfg
