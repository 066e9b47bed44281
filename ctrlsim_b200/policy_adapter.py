"""Reference-signature policy: the B200 hot path behind the exact per-scene ``Policy`` interface of the reference.

``B200Policy`` / ``B200PolicyEvaluator`` (evaluator.py) batch whole scene sets on the GPU and replace the reference's
evaluator loop.  This module is the other half of the drop-in boundary: a class with the reference's own
constructor and method signatures,

    AutoregressivePolicy(cfg, model_path, model, use_rtg, predict_rtgs, discretize_rtgs, real_time_rewards,
                         privileged_return, max_return, min_return, key_dict, tilt_dict, name, action_temperature,
                         nucleus_sampling, nucleus_threshold)                    policies/autoregressive_policy.py:10-48
    reset(vehicle_data_dict)                                                       policies/policy.py:45-59
    update_state(vehicle_data_dict, vehicles_to_evaluate, t)                       policies/policy.py:68-105
    predict(vehicle_data_dict, gt_data_dict, preproc_data, dset, vehicles_to_evaluate, t) -> vehicle_data_dict
                                                                                   policies/autoregressive_policy.py:168-253
    act(veh, t, vehicle_data_dict) -> (veh, [acceleration, steering])              policies/autoregressive_policy.py:256-274

so that the STOCK ``PolicyEvaluator`` (evaluators/policy_evaluator.py:426-595: its own Nocturne simulator, its own
``vehicle_data_dict``) can drive the CUDA policy one scene at a time: ``PolicyEvaluator(cfg, B200AutoregressivePolicy(...))``.
Per step the adapter mirrors the dict state into the one-scene ``CtrlSimBatch`` (history of states and applied actions),
runs focal grouping -> tokenisation -> two-pass network -> RTG / action sampling through the C-ABI
(``ctrlsim_plan_groups`` + ``ctrlsim_policy_step``) and writes ``next_acceleration`` / ``next_steering`` / ``rtgs`` back
through ``key_dict``.  The simulator, rewards and metrics stay the evaluator's.  Throughput-wise this is the slow way to
use the library (one scene, a device round trip per step); it exists for drop-in use and for parity checks against the
reference's own loop.
"""
from __future__ import annotations

import numpy as np

from .evaluator import B200Policy


class _DeviceBackend:
    """ctrlsim_plan_groups + ctrlsim_policy_step on a one-scene batch (what B200Policy.predict does)."""

    def __init__(self, inner: B200Policy):
        self.inner = inner

    def make_batch(self, cfg, scene, scene_id, parsed, evaluated):
        from .batch import SceneBatch
        b = SceneBatch(cfg, [scene], [scene_id], self.inner.model.device, eval_threshold=len(evaluated), parsed=[parsed],
                       evaluated_sets=[evaluated])
        b.reset_dynamic()
        self.inner.attach_caches(b)
        return b

    def step(self, batch, t, states_t, actions_prev, rtgs_t=None):
        import torch
        dev = batch.device
        batch.t["hist_state"][0, :, t] = torch.from_numpy(states_t).to(dev)
        if actions_prev is not None:
            batch.t["hist_action"][0, :, t - 1] = torch.from_numpy(actions_prev).to(dev)
        if rtgs_t is not None:  # real_time_rewards: the evaluator's tracked, un-normalised RTGs of step t
            batch.t["rt_rtg"][0, :, t] = torch.from_numpy(np.ascontiguousarray(rtgs_t, np.float64)).to(dev)
        self.inner.predict(batch, t)
        torch.cuda.synchronize(dev)
        return (batch.t["next_action"][0].cpu().numpy(), batch.t["tr_rtg_idx"][0, :, t].cpu().numpy(),
                batch.t["tr_act_idx"][0, :, t].cpu().numpy())


class B200AutoregressivePolicy:
    def __init__(self, cfg, model_path, model, use_rtg, predict_rtgs, discretize_rtgs, real_time_rewards,
                 privileged_return, max_return, min_return, key_dict, tilt_dict, name, action_temperature,
                 nucleus_sampling, nucleus_threshold, seed=0, backend=None):
        """Arguments as in the reference.  ``model``: a ``DeviceModel``.  ``seed``: key of the explicit sampler (DESIGN
        section 4).  ``backend`` (tests): object with ``make_batch`` / ``step`` standing in for the device."""
        # real_time_rewards policies (cfgs/policy/dt.yaml): the stock evaluator computes the dense reward and tracks the
        # RTGs itself (policy_evaluator.py:123-149, evaluator.py:106-140); update_state() picks the RTGs of step t up from
        # vehicle_data_dict like policies/policy.py:91-93 and predict() hands them to the device (CtrlSimBatch.rt_rtg)
        if backend is None:
            self.inner = B200Policy(cfg, model_path, model, use_rtg, predict_rtgs, discretize_rtgs, real_time_rewards,
                                    privileged_return, max_return, min_return, key_dict, tilt_dict, name,
                                    action_temperature, nucleus_sampling, nucleus_threshold, seed=seed)
            backend = _DeviceBackend(self.inner)
        elif not predict_rtgs and not real_time_rewards:
            raise NotImplementedError("RTGs that are neither predicted nor tracked are not implemented")
        self.backend = backend
        # what the reference's Policy.__init__ keeps and its evaluator reads (policy_evaluator.py:39-41,123-143,417-423)
        self.cfg = cfg.copy()
        self.model_path, self.model, self.name = model_path, model, name
        self.cfg_model, self.cfg_rl_waymo = cfg.model, cfg.dataset.waymo
        self.steps = self.cfg.nocturne.steps
        self.use_rtg, self.predict_rtgs, self.discretize_rtgs = use_rtg, predict_rtgs, discretize_rtgs
        self.real_time_rewards, self.privileged_return = real_time_rewards, privileged_return
        self.max_return, self.min_return = max_return, min_return
        self.key_dict, self.tilt_dict = key_dict, tilt_dict
        self.action_temperature = action_temperature
        self.nucleus_sampling, self.nucleus_threshold = nucleus_sampling, nucleus_threshold
        if tilt_dict["tilt"]:
            self.goal_tilt, self.veh_veh_tilt, self.veh_edge_tilt = (tilt_dict["goal_tilt"], tilt_dict["veh_veh_tilt"],
                                                                     tilt_dict["veh_edge_tilt"])
        self.scene_index = -1  # sampler counter word 0; reset() advances it (set it before reset() to pin a scene id)
        self._batch = None

    # ---- Policy.reset ---------------------------------------------------------------------------------------------
    def reset(self, vehicle_data_dict):
        n = len(vehicle_data_dict.keys())
        self.states = np.zeros((n, self.steps, 8))
        self.actions = np.zeros((n, self.steps, 2))
        self.rtgs = np.zeros((n, self.steps, 3))
        self.idx_to_veh_id, self.veh_id_to_idx = {}, {}
        for i, v in enumerate(vehicle_data_dict.keys()):
            self.idx_to_veh_id[i] = v
            self.veh_id_to_idx[v] = i
        self.scene_index += 1
        self._batch = None

    # ---- Policy.update_state ----------------------------------------------------------------------------------------
    def update_state(self, vehicle_data_dict, vehicles_to_evaluate, t):
        for i, v in enumerate(vehicle_data_dict.keys()):
            d = vehicle_data_dict[v]
            self.states[i, t] = (d["position"][t]["x"], d["position"][t]["y"], d["velocity"][t]["x"], d["velocity"][t]["y"],
                                 d["heading"][t], d["length"], d["width"], d["existence"][t])
            if t > 0:
                self.actions[i, t - 1] = (d["acceleration"][t - 1], d["steering"][t - 1])
            if self.real_time_rewards and self.use_rtg:  # policies/policy.py:91-93
                self.rtgs[i, t] = np.asarray(d[self.key_dict["rtgs"]][t], np.float64)

    # ---- AutoregressivePolicy.predict -------------------------------------------------------------------------------
    def _make_batch(self, vehicle_data_dict, gt_data_dict, preproc_data, vehicles_to_evaluate):
        ids = list(vehicle_data_dict.keys())
        n, T1 = len(ids), self.steps + 1
        gt = np.zeros((n, T1, 4), np.float32)
        gt_valid = np.zeros((n, T1), np.uint8)
        size = np.zeros((n, 2), np.float32)
        goal = np.zeros((n, 4), np.float64)
        for i, v in enumerate(ids):
            traj = np.asarray(gt_data_dict[v]["traj"], np.float64)  # [steps + 1, 8]: x, y, heading, speed, existence, ...
            gt[i], gt_valid[i] = traj[:T1, :4], traj[:T1, 4] != 0
            d = vehicle_data_dict[v]
            size[i] = (d["length"], d["width"])
            goal[i] = (d["goal_position"]["x"], d["goal_position"]["y"], d["goal_heading"], d["goal_speed"])
        parsed = dict(n=n, gt=gt, gt_valid=gt_valid, size=size, moving=np.ones(n, bool), goal=goal,
                      goal_norm=np.zeros(n), segs=np.zeros((0, 4), np.float32))
        evaluated = [self.veh_id_to_idx[v] for v in vehicles_to_evaluate]
        scene = {"preproc": {"road_points": preproc_data["road_points"], "road_types": preproc_data["road_types"]}}
        return self.backend.make_batch(self.cfg, scene, self.scene_index, parsed, evaluated)

    def predict(self, vehicle_data_dict, gt_data_dict, preproc_data, dset, vehicles_to_evaluate, t):
        if self._batch is None:
            if t != 0:
                raise RuntimeError("predict() must be called for every step from t = 0 on after reset()")
            self._batch = self._make_batch(vehicle_data_dict, gt_data_dict, preproc_data, vehicles_to_evaluate)
            self._evaluated = {self.veh_id_to_idx[v] for v in vehicles_to_evaluate}
        if self.real_time_rewards:
            next_action, rtg_idx, act_idx = self.backend.step(self._batch, t, self.states[:, t],
                                                              self.actions[:, t - 1] if t > 0 else None, self.rtgs[:, t])
        else:
            next_action, rtg_idx, act_idx = self.backend.step(self._batch, t, self.states[:, t],
                                                              self.actions[:, t - 1] if t > 0 else None)
        self.last_rtg_idx, self.last_act_idx = rtg_idx, act_idx  # sampled bins of this step (-1: none), for parity dumps
        w, kd = self.cfg_rl_waymo, self.key_dict
        R = w.rtg_discretization - 1
        lo = (w.min_rtg_pos, w.min_rtg_veh, w.min_rtg_road)
        hi = (w.max_rtg_pos, w.max_rtg_veh, w.max_rtg_road)
        for i, v in self.idx_to_veh_id.items():
            d = vehicle_data_dict[v]
            if not self.predict_rtgs:  # the evaluator owns the RTG series (autoregressive_policy.py:243-248 is skipped)
                pass
            elif rtg_idx[i, 0] >= 0:  # an RTG was drawn for this vehicle this step (it is in some focal group's context)
                rtg = np.array([rtg_idx[i, c] / R * (hi[c] - lo[c]) + lo[c] for c in range(3)])  # undiscretize_rtgs
                d["next_rtg_goal"], d["next_rtg_veh"], d["next_rtg_road"] = rtg
                d[kd["rtgs"]].append(rtg)
            else:
                d[kd["rtgs"]].append(np.array([0] * self.cfg_model.num_reward_components))
            if i in self._evaluated:
                if act_idx[i] >= 0:
                    d[kd["next_acceleration"]], d[kd["next_steering"]] = float(next_action[i, 0]), float(next_action[i, 1])
                elif not d["existence"][t]:  # a focal vehicle that no longer exists (autoregressive_policy.py:249-251)
                    d[kd["next_acceleration"]], d[kd["next_steering"]] = 0.0, 0.0
        return vehicle_data_dict

    # ---- AutoregressivePolicy.act -----------------------------------------------------------------------------------
    def act(self, veh, t, vehicle_data_dict):
        d = vehicle_data_dict[veh.getID()]
        if not d["existence"][-1]:
            acceleration = steering = 0.0
            veh.setPosition(-1000000, -1000000)  # the reference parks vehicles that ran out of states far away
        else:
            acceleration, steering = d[self.key_dict["next_acceleration"]], d[self.key_dict["next_steering"]]
        if acceleration > 0.0:
            veh.acceleration = acceleration
        else:
            veh.brake(np.abs(acceleration))
        veh.steering = steering
        return veh, [acceleration, steering]
