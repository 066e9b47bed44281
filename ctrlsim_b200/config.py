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


def dt_config(wide: bool = False) -> AttrDict:
    """The decision-transformer baseline: cfgs/model/dt.yaml on top of cfgs/model/ctrl_sim.yaml
    (continuous RTG inputs, (rtg, state, action) token order, no RTG head, no future-state head) and cfgs/policy/dt.yaml
    (RTGs are not predicted but tracked in real time: they start at the maximum return and are decremented by the dense
    reward of every step)."""
    cfg = default_config(wide)
    cfg.model.decision_transformer, cfg.model.predict_rtg, cfg.model.predict_future_states = True, False, False
    cfg.eval.policy.update(run_name="dt", model="dt", use_rtg=True, predict_rtgs=False, discretize_rtgs=False,
                           real_time_rewards=True, max_return=True)
    return cfg

