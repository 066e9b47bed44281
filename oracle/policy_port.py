"""ORACLE (test infrastructure only): numpy restatement of the reference's closed-loop evaluation of one scene.

Rows of SURVEY 8(a) restated here (reference file:line):
  T1  Policy.reset / update_state                       policies/policy.py:45-59,68-105
  T2  AutoregressivePolicy.get_data                      policies/autoregressive_policy.py:51-165
      select_relevant_agents                             datasets/rl_waymo/dataset.py:278-319
  T3  discretize_actions / discretize_rtgs               datasets/rl_waymo/dataset.py:365-387
      normalize_scene, apply_se2_transform, angle_sub    datasets/rl_waymo/dataset.py:390-428, utils/geometry.py:3-47
  M8  process_predicted_rtg, undiscretize_rtgs           policies/policy.py:108-142, dataset.py:351-362
  M9  action sampling, undiscretize_actions, act         policies/autoregressive_policy.py:168-274, dataset.py:322-339
  S5  compute_reward, update_vehicle_data_dict, goals    utils/sim.py:83-141, evaluators/policy_evaluator.py:99-159,
                                                         evaluators/evaluator.py:60-104
  S6  apply_gt_action, BicycleModel.backward             evaluators/evaluator.py:160-193, nocturne/bicycle_model.py:51-109
  S7  update_running_statistics, compute_metrics         evaluators/policy_evaluator.py:162-305
  loop evaluate_policy                                   evaluators/policy_evaluator.py:426-595
The simulator is oracle/sim_port.py (C restatement, incl. Box2D's vehicle-vehicle contact response) and the network is oracle/model_port.py; sampling uses the
explicit sampler contract of oracle/sampler.py.  Pinned against the reference itself by the fixtures under
tests/golden/ (oracle/make_golden.py).

Deliberately preserved quirks: the served-vehicle loop removes from the list it iterates (so every other candidate is
skipped, autoregressive_policy.py:123-127); np.round half-to-even; yaw sign convention of normalize_scene; angular
speed histogram uses heading/dt; RTG (0,0,0) appended for vehicles in no context; with real_time_rewards the dense
reward's goal / collision terms are those of step 0 (see run_scene).
"""
from __future__ import annotations

import math
import random

import numpy as np
import torch

from . import sampler, sim_port

TWO_PI = 2 * np.pi


def angle_sub(cur, tgt):
    d = (tgt - cur) % TWO_PI
    if d > np.pi:
        d = -(TWO_PI - d)
    return d


def angle_sub_arr(cur, tgt):
    d = (tgt - cur) % TWO_PI
    d = np.where(d > np.pi, -(TWO_PI - d), d)
    return d


def se2(coords, translation, yaw):
    c = coords - translation
    R = np.array([[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]])
    shp = c.shape
    return (R @ c.reshape(-1, 2).T).T.reshape(shp)


def inverse_bicycle(gt_next, sim_pos, sim_theta, sim_vel, dt):
    """BicycleModel(x,y,theta,vel,L of GT t+1).backward(prev = simulated state) -> (accel, steer)."""
    vel_gt, theta_gt, L = gt_next[3], gt_next[2], gt_next[7]
    accel = (vel_gt - sim_vel) / dt
    w = angle_sub(sim_theta, theta_gt) / dt
    C = 2.0 * L * w / (vel_gt + sim_vel + 1e-10)
    with np.errstate(invalid="ignore"):
        steer = np.arctan(2.0 * C / np.sqrt(4 - C ** 2))
    if np.isnan(steer):
        steer = 0.0
    return float(accel), float(np.clip(steer, -0.7, 0.7))


class MetricsPort:
    """S7 accumulators + compute_metrics."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.w = cfg.dataset.waymo
        self.steps, self.dt, self.hist = cfg.nocturne.steps, cfg.nocturne.dt, cfg.nocturne.history_steps
        self.goal, self.ade, self.fde, self.coll, self.off = [], [], [], [], []
        self.samples = {k: [] for k in ("lin_sim", "lin_gt", "ang_sim", "ang_gt", "acc_sim", "acc_gt", "nd_sim", "nd_gt")}

    def add_scene(self, rec, evaluated):
        T = self.steps + 1
        colls, offs = [], []
        for v in evaluated:
            mask = rec["existence"][v].astype(bool).copy()
            mask[: self.hist] = False
            if mask.sum() == 0:
                continue
            rew = rec["reward"][v][mask]
            self.goal.append(float(np.any(rew[:, 0] == 1)))
            colls.append(float(np.any(rew[:, 6] == 1)))
            offs.append(float(np.any(rew[:, 7] == 1)))
            d = np.linalg.norm(rec["pos"][v][mask] - rec["gt_pos"][v][mask], axis=1)
            self.ade.append(d.mean())
            last = np.where(mask)[0][-1]
            self.fde.append(np.linalg.norm(rec["pos"][v][last] - rec["gt_pos"][v][last]))
            self.samples["lin_sim"].append(np.linalg.norm(rec["vel"][v][mask], axis=1))
            self.samples["lin_gt"].append(rec["gt_speed"][v][mask])
            self.samples["ang_sim"].append(rec["heading"][v][mask] / self.dt)
            self.samples["ang_gt"].append(rec["gt_heading"][v][mask] / self.dt)
            am = np.ones(int(mask.sum()), bool)
            am[0] = am[-1] = False
            self.samples["acc_gt"].append(rec["gt_accel"][v][mask][am])
            self.samples["acc_sim"].append(rec["accel"][v][mask][am])
            self.samples["nd_gt"].append(rec["gt_nearest_dist"][v][mask])
            self.samples["nd_sim"].append(rec["nearest_dist"][v][mask])
        if colls:
            self.coll.append(np.mean(colls))
            self.off.append(np.mean(offs))

    @staticmethod
    def jsd(p, q):
        """scipy.spatial.distance.jensenshannon (base e): sqrt(0.5*(KL(p||m)+KL(q||m))) after normalising p, q."""
        p = np.asarray(p, np.float64)
        q = np.asarray(q, np.float64)
        p, q = p / p.sum(), q / q.sum()
        m = (p + q) / 2.0

        def rel(a, b):
            out = np.zeros_like(a)
            nz = a > 0
            out[nz] = a[nz] * np.log(a[nz] / b[nz])
            return out
        return float(np.sqrt((rel(p, m).sum() + rel(q, m).sum()) / 2.0))

    def compute(self):
        w = self.w
        cat = {k: (np.concatenate(v) if v else np.zeros(0)) for k, v in self.samples.items()}
        out = {"goal": float(np.mean(self.goal)), "collision_rate": float(np.mean(self.coll)),
               "offroad_rate": float(np.mean(self.off)), "fde": float(np.mean(self.fde)),
               "ade": float(np.mean(self.ade))}

        def hist_jsd(sim, gt, edges):
            P = np.histogram(sim, bins=edges)[0] / len(sim)
            Q = np.histogram(gt, bins=edges)[0] / len(gt)
            return self.jsd(P, Q)
        out["lin_speed_jsd"] = hist_jsd(np.clip(cat["lin_sim"], 0, 30), np.clip(cat["lin_gt"], 0, 30),
                                        np.arange(201) * 0.5 * (100 / 30))
        out["ang_speed_jsd"] = hist_jsd(np.clip(cat["ang_sim"], -50, 50), np.clip(cat["ang_gt"], -50, 50),
                                        np.arange(201) * 0.5 - 50)
        g = (np.clip(cat["acc_gt"], w.min_accel, w.max_accel) - w.min_accel) / (w.max_accel - w.min_accel)
        g = np.round(g * (w.accel_discretization - 1)) / (w.accel_discretization - 1)
        g = g * (w.max_accel - w.min_accel) + w.min_accel
        out["accel_jsd"] = hist_jsd(cat["acc_sim"], g, np.arange(w.accel_discretization + 1) * 2 - w.accel_discretization)
        out["nearest_dist_jsd"] = hist_jsd(np.clip(cat["nd_sim"], 0, 40), np.clip(cat["nd_gt"], 0, 40),
                                           np.arange(201) * 0.5 * (100 / 40))
        return out


class RolloutPort:
    def __init__(self, cfg, model, seed=0, tilts=(0, 0, 0), temperature=1.0, eval_threshold=None, nucleus=None,
                 contacts=True, fp64_trig=False, predict_rtgs=True, discretize_rtgs=True, real_time_rewards=False,
                 max_return=False, min_return=False, use_rtg=True):
        self.cfg, self.model = cfg, model
        # Policy constructor switches (policies/policy.py:9-39; cfgs/policy/dt.yaml sets predict_rtgs=False,
        # discretize_rtgs=False, real_time_rewards=True, max_return=True): with real_time_rewards the RTG series is the
        # un-normalised return still to collect, decremented by the dense reward of every step
        # (policy_evaluator.py:123-149) instead of being sampled from the RTG head
        self.predict_rtgs, self.discretize_rtgs, self.real_time = predict_rtgs, discretize_rtgs, real_time_rewards
        self.max_return, self.min_return = max_return, min_return
        # use_rtg=False: Policy.update_state never copies RTGs into the policy's buffers (policies/policy.py:89-95), the
        # network is fed RTG (0, 0, 0); this is what cfgs/policy/dt.yaml composes to as shipped (its key is `use_rtgs`)
        self.use_rtg = use_rtg
        self.w, self.m = cfg.dataset.waymo, cfg.model
        self.seed, self.tilts, self.temperature = seed, tilts, float(temperature)
        self.nucleus = nucleus  # None, or the nucleus threshold p (autoregressive_policy.py:216-230)
        self.contacts = contacts  # simulator with Box2D contact response (sim_oracle.c second half) or contact-free
        self.fp64_trig = fp64_trig  # simulator trig rounded like the GPU path's (sim_oracle.c SIMO_TRIG_FP64)
        self.steps, self.dt, self.hist = cfg.nocturne.steps, cfg.nocturne.dt, cfg.nocturne.history_steps
        self.eval_threshold = eval_threshold if eval_threshold is not None else cfg.eval.multi_agent_eval_threshold
        self.metrics = MetricsPort(cfg)
        random.seed(cfg.eval.seed)  # PolicyEvaluator.reset (policy_evaluator.py:45-50)
        self.n_forwards = 0

    # ------------------------------------------------------------------------------------------------ T2/T3
    def plan_groups(self, ep, t):
        """Greedy focal grouping. Returns [(focal, closest_ids(sorted), served)], dead list."""
        w = self.w
        T = w.train_context_length
        t0 = 0 if t < T else t - (T - 1)
        unacc = list(ep["eval_order"])
        groups, dead = [], []
        while unacc:
            focal = unacc.pop(0)
            if not ep["states"][focal, t, -1]:
                dead.append(focal)
                continue
            if len(ep["road_points"]) == 0:  # scene without road polylines (autoregressive_policy.py:106-108)
                dead.append(focal)
                continue
            if t == 0:
                ep["relevant"][focal] = []
            dist = np.linalg.norm(ep["states"][focal, t0, :2][None] - ep["states"][:, t0, :2], axis=-1)
            valid = np.where(dist < w.agent_dist_threshold)[0]
            rel = ep["relevant"][focal]
            if len(rel) == 0:
                closest = np.intersect1d(np.argsort(dist)[: w.max_num_agents], valid)
            else:
                closest = np.intersect1d(np.array(rel).astype(int), valid)
                if len(closest) < len(rel):
                    rel = [i for i in rel if i in closest]
            served = [focal]
            i = 0
            while i < len(unacc):  # list mutated while iterated: the element after a removed one is skipped
                u = unacc[i]
                if u in closest:
                    served.append(u)
                    unacc.remove(u)
                i += 1
            new_rel = [int(i) for i in closest] if t == 0 else rel
            for v in served:
                ep["relevant"][v] = new_rel
            groups.append((int(focal), closest.astype(int), served, list(new_rel)))
        return groups, dead

    def tokenize(self, ep, t, focal, closest):
        """One focal group's model inputs (float64 numpy, as the reference builds them)."""
        w = self.w
        T, A, P = w.train_context_length, w.max_num_agents, w.max_num_road_polylines
        sl = slice(0, T) if t < T else slice(t - (T - 1), t + 1)
        states = ep["states"][:, sl]
        actions = ep["actions"][:, sl]
        rtgs = ep["rtgs"][:, sl].copy()
        lo = (w.min_rtg_pos, w.min_rtg_veh, w.min_rtg_road)
        hi = (w.max_rtg_pos, w.max_rtg_veh, w.max_rtg_road)
        for c in range(3):
            rtgs[:, :, c] = (np.clip(rtgs[:, :, c], lo[c], hi[c]) - lo[c]) / (hi[c] - lo[c])
        goals = ep["goals"][:, sl][:, 0]
        timesteps = ep["timesteps"][0, sl, 0].astype(int)
        n = len(closest)
        st = np.zeros((A, T, 8))
        ty = -np.ones((A, 5))
        ac = np.zeros((A, T, 2))
        rt = np.zeros((A, T, 3))
        go = np.zeros((A, w.goal_dim))
        st[:n], ty[:n], ac[:n], rt[:n], go[:n] = states[closest], ep["types"][closest], actions[closest], rtgs[closest], goals[closest]
        origin = int(np.where(closest == focal)[0][0])
        # discretize (half-to-even)
        a0 = (np.clip(ac[:, :, 0], w.min_accel, w.max_accel) - w.min_accel) / (w.max_accel - w.min_accel)
        a1 = (np.clip(ac[:, :, 1], w.min_steer, w.max_steer) - w.min_steer) / (w.max_steer - w.min_steer)
        act_idx = np.round(a0 * (w.accel_discretization - 1)) * w.steer_discretization + np.round(a1 * (w.steer_discretization - 1))
        rtg_idx = np.round(rt * (w.rtg_discretization - 1)) if self.discretize_rtgs else rt  # autoregressive_policy.py:141-142
        # normalize_scene
        yaw = st[origin, 0, 4]
        rot = (np.pi / 2) + np.sign(-yaw) * np.abs(yaw)
        trans = st[origin, 0, :2].copy()
        st[:, :, :2] = se2(st[:, :, :2], trans[None, None], rot)
        st[:, :, 2:4] = se2(st[:, :, 2:4], np.zeros((1, 1, 2)), rot)
        st[:, :, 4] = angle_sub_arr(st[:, :, 4], -rot)
        go[:, :2] = se2(go[:, :2], trans[None], rot)
        go[:, 2:4] = se2(go[:, 2:4], np.zeros((1, 2)), rot)
        go[:, 4] = angle_sub_arr(go[:, 4], -rot)
        rp = ep["road_points"].copy()
        rp[:, :, :2] = se2(rp[:, :, :2], trans[None, None], rot)
        if len(rp) > P:
            dmax = (np.linalg.norm(rp[:, :, :2], axis=-1) * rp[:, :, -1]).max(1)
            keep = np.argsort(dmax)[:P]
            frp, frt = rp[keep], ep["road_types"][keep]
        else:
            frp = np.zeros((P,) + rp.shape[1:])
            frp[: len(rp)] = rp
            frt = -np.ones((P, ep["road_types"].shape[1]))
            frt[: len(rp)] = ep["road_types"]
        return {"agent_states": st, "agent_types": ty, "goals": go, "actions": act_idx, "rtgs": rtg_idx,
                "timesteps": np.repeat(timesteps[None, :, None], A, axis=0), "road_points": frp, "road_types": frt}

    def _forward(self, tok):
        self.n_forwards += 1
        data = {k: torch.from_numpy(np.ascontiguousarray(v))[None] for k, v in tok.items()}
        return self.model.forward(data)

    # ------------------------------------------------------------------------------------------------ loop
    def setup_scene(self, scene_idx, scen_json):
        """Simulator, ground truth, goals and the empty per-scene record (evaluator.py:60-84, policy_evaluator.py:69-96)."""
        steps, dt = self.steps, self.dt
        parsed = sim_port.parse_scenario(scen_json)
        n = parsed["n"]
        gt = sim_port.ground_truth(parsed, steps)
        sim = sim_port.ScenePort(parsed, contacts=self.contacts, fp64_trig=self.fp64_trig)
        goal = np.zeros((n, 4))
        for i in range(n):
            gp, gh, gs = parsed["target"][i, :2].astype(np.float64), float(parsed["target"][i, 2]), float(parsed["target"][i, 3])
            gone = np.where(gt[i, :, 4] == 0)[0]
            if len(gone) > 0:
                k = gone[0] - 1
                if np.linalg.norm(gt[i, k, :2] - gp) > 0.0:
                    gp, gh, gs = gt[i, k, :2], gt[i, k, 2], gt[i, k, 3]
            goal[i] = (gp[0], gp[1], gh, gs)
        normalizer = np.linalg.norm(sim.position() - goal[:, :2], axis=1)
        T1 = steps + 1
        rec = {"scene": scene_idx, "n": n, "pos": np.zeros((n, T1, 2)), "vel": np.zeros((n, T1, 2)),
               "heading": np.zeros((n, T1)), "existence": np.zeros((n, T1)), "accel": np.zeros((n, T1)),
               "steer": np.zeros((n, T1)), "reward": np.zeros((n, T1, 8)),
               "nearest_dist": np.zeros((n, T1)), "gt_nearest_dist": np.zeros((n, T1)),
               "gt_pos": gt[:, :, :2].copy(), "gt_heading": gt[:, :, 2].copy(), "gt_speed": gt[:, :, 3].copy(),
               "gt_accel": np.zeros((n, T1)), "goal": goal, "size": parsed["size"].astype(np.float64)}
        rec["gt_accel"][:, 1:steps - 1] = (gt[:, 2:steps, 3] - gt[:, 0:steps - 2, 3]) / (2 * dt)
        return {"parsed": parsed, "n": n, "gt": gt, "sim": sim, "goal": goal, "normalizer": normalizer, "rec": rec}

    def policy_record(self, n):
        """What one Policy object adds to the scene record: sampled bins, RTGs (its key_dict['rtgs']), focal groups."""
        steps = self.steps
        return {"rtgs": np.zeros((n, steps, 3)), "rtg_idx": -np.ones((steps, n, 3), np.int32),
                "act_idx": -np.ones((steps, n), np.int32), "groups": [[] for _ in range(steps)], "logits": {}}

    def new_episode(self, ctx, evaluated, preproc):
        """Policy.reset (policies/policy.py:45-59) + the served order of get_data (autoregressive_policy.py:88-94)."""
        n, steps, w, gt = ctx["n"], self.steps, self.w, ctx["gt"]
        lengths = [int(gt[v][:, 4].sum()) for v in evaluated]
        order = np.argsort(np.array(lengths))[::-1]
        return {
            "states": np.zeros((n, steps, 8)), "actions": np.zeros((n, steps, 2)), "rtgs": np.zeros((n, steps, 3)),
            "goals": np.zeros((n, steps, w.goal_dim)), "timesteps": np.zeros((n, steps, 1)),
            "types": np.tile(np.eye(5)[1], (n, 1)), "relevant": {},
            "eval_order": list(np.array(evaluated)[order]),
            "road_points": preproc["road_points"], "road_types": preproc["road_types"],
        }

    def observe(self, ctx, t):
        """update_vehicle_data_dict (policy_evaluator.py:99-159) incl. compute_reward (utils/sim.py:83-141)."""
        sim, rec, gt, goal, normalizer, n = ctx["sim"], ctx["rec"], ctx["gt"], ctx["goal"], ctx["normalizer"], ctx["n"]
        rew_cfg = self.cfg.nocturne["rew_cfg"]
        pos, head, spd, vel = sim.position(), sim.heading(), sim.speed(), sim.velocity()
        cv, ce = sim.collisions()
        rec["pos"][:, t], rec["vel"][:, t], rec["heading"][:, t] = pos, vel, head
        ex = gt[:, t, 4].copy()
        if t > 0:
            ex[rec["existence"][:, t - 1] == 0] = 0
        rec["existence"][:, t] = ex
        for i in range(n):
            prev = t > 0 and rec["reward"][i, t - 1, 0]
            dist = np.linalg.norm(goal[i, :2] - pos[i])
            r0 = 1.0 if prev else float(dist < rew_cfg["position_target_tolerance"])
            r2 = float(np.abs(goal[i, 3] - spd[i]) < rew_cfg["speed_target_tolerance"])
            r1 = float(np.abs(angle_sub(goal[i, 2], head[i])) < rew_cfg["heading_target_tolerance"])
            gds, rs = rew_cfg["shaped_goal_distance_scaling"], rew_cfg["reward_scaling"]
            nz = normalizer[i] if normalizer[i] != 0.0 else 1.0
            r3 = gds / rs if prev else gds * (1 - dist / nz) / rs
            r4 = gds * (1 - np.abs(spd[i] - goal[i, 3]) / 40.0) / rs
            r5 = gds * (1 - np.abs(angle_sub(head[i], goal[i, 2])) / (2 * np.pi)) / rs
            rec["reward"][i, t] = (r0, r1, r2, r3, r4, r5, float(cv[i]), float(ce[i]))
        for key, P in (("nearest_dist", pos), ("gt_nearest_dist", gt[:, t, :2])):
            Pm = np.where(ex[:, None].astype(bool), P, np.inf)
            with np.errstate(invalid="ignore"):
                d2 = ((Pm[:, None] - Pm[None]) ** 2).sum(-1)
            np.fill_diagonal(d2, np.inf)
            with np.errstate(invalid="ignore"):
                d = np.sqrt(np.min(d2, axis=1))
            d[d == np.inf] = np.nan
            rec[key][:, t] = np.nan_to_num(d * ex, nan=0.0) * ex

    def update_state(self, ep, ctx, prec, t):
        """Policy.update_state (policies/policy.py:68-105): world record -> this policy's buffers; RTGs from its own key."""
        rec, goal, w = ctx["rec"], ctx["goal"], self.w
        ep["states"][:, t, :2], ep["states"][:, t, 2:4] = rec["pos"][:, t], rec["vel"][:, t]
        ep["states"][:, t, 4], ep["states"][:, t, 5:7] = rec["heading"][:, t], ctx["parsed"]["size"]
        ep["states"][:, t, 7] = rec["existence"][:, t]
        ep["timesteps"][:, t, 0] = t
        if t > 0:
            ep["actions"][:, t - 1, 0], ep["actions"][:, t - 1, 1] = rec["accel"][:, t - 1], rec["steer"][:, t - 1]
            if self.use_rtg:
                ep["rtgs"][:, t - 1] = prec["rtgs"][:, t - 1]
        if self.real_time and self.use_rtg:  # policies/policy.py:93-95: the RTG of the CURRENT step is known before the forward
            ep["rtgs"][:, t] = prec["rtgs"][:, t]
        ep["goals"][:, t] = np.stack([goal[:, 0], goal[:, 1], goal[:, 3] * np.cos(goal[:, 2]),
                                      goal[:, 3] * np.sin(goal[:, 2]), goal[:, 2]], -1)[:, : w.goal_dim]

    def predict_step(self, ep, prec, t, scene_idx, next_act, logit_steps=()):
        """AutoregressivePolicy.predict (autoregressive_policy.py:168-253) for this policy's served vehicles: writes the
        sampled bins / RTGs / groups into ``prec`` and the continuous actions into ``next_act`` (its key_dict slots)."""
        w = self.w
        groups, dead = self.plan_groups(ep, t)
        ti = t if t < w.train_context_length else w.train_context_length - 1
        done = {}
        for g, (focal, closest, served, rel) in enumerate(groups):
            members = -np.ones(w.max_num_agents, np.int32)
            members[: len(closest)] = closest
            prec["groups"][t].append({"focal": focal, "members": members, "served": [int(v) for v in served]})
            slot = {int(a): k for k, a in enumerate(closest)}
            tok = self.tokenize(ep, t, focal, closest)
            if self.predict_rtgs:
                out = self._forward(tok)
                rtg_logits = out["rtg_preds"][0, :, ti].numpy()
                for a in rel:
                    if a not in done:
                        tilted = a in served
                        lg = rtg_logits[slot[a]].reshape(w.rtg_discretization, 3)
                        done[a] = [sampler.sample_from_x(sampler.rtg_x(lg[:, c], self.tilts[c] if tilted else 0),
                                                         self.seed, scene_idx, a, t, c) for c in range(3)]
                        prec["rtg_idx"][t, a] = done[a]
                    tok["rtgs"][slot[a], ti] = done[a]
            out2 = self._forward(tok)  # the only forward when the RTGs are not predicted (autoregressive_policy.py:210)
            act_logits = out2["action_preds"][0, :, ti].numpy()
            if t in logit_steps:
                prec["logits"][(t, g)] = {"action_logits": act_logits.copy()}
                if self.predict_rtgs:
                    prec["logits"][(t, g)]["rtg_logits"] = rtg_logits.copy()
            for v in served:
                ax = sampler.action_x(act_logits[slot[v]], self.temperature)
                if self.nucleus is None:
                    idx = sampler.sample_from_x(ax, self.seed, scene_idx, v, t, sampler.COMP_ACTION)
                else:
                    idx = sampler.sample_from_x_nucleus(ax, self.nucleus, self.seed, scene_idx, v, t, sampler.COMP_ACTION)
                prec["act_idx"][t, v] = idx
                next_act[v, 0] = (idx // w.steer_discretization) / (w.accel_discretization - 1) * (w.max_accel - w.min_accel) + w.min_accel
                next_act[v, 1] = (idx % w.steer_discretization) / (w.steer_discretization - 1) * (w.max_steer - w.min_steer) + w.min_steer
        R = w.rtg_discretization - 1
        for a, idx in done.items():
            prec["rtgs"][a, t] = (idx[0] / R * (w.max_rtg_pos - w.min_rtg_pos) + w.min_rtg_pos,
                                  idx[1] / R * (w.max_rtg_veh - w.min_rtg_veh) + w.min_rtg_veh,
                                  idx[2] / R * (w.max_rtg_road - w.min_rtg_road) + w.min_rtg_road)
        for v in dead:
            next_act[v] = 0.0

    def apply_controls(self, ctx, t, controlled, next_act, targets=None):
        """The per-vehicle loop of evaluate_policy (policy_evaluator.py:526-535): policy.act for controlled vehicles
        from step history_steps - 1 on, apply_gt_action (evaluator.py:160-193) otherwise.  ``targets`` optionally
        replaces the log-replay target states of some vehicles ({vehicle: [T1, 8] array})."""
        sim, rec, gt, n, dt = ctx["sim"], ctx["rec"], ctx["gt"], ctx["n"], self.dt
        pos, head, spd = sim.position(), sim.heading(), sim.speed()
        for i in range(n):
            if t >= self.hist - 1 and i in controlled:
                if not rec["existence"][i, t]:
                    a, s = 0.0, 0.0
                    sim.teleport(i, -1000000, -1000000)
                else:
                    a, s = next_act[i]
            else:
                exists = gt[i, t, 4] and gt[i, t + 1, 4]
                if t > 0 and rec["existence"][i, t] == 0:
                    exists = 0
                if not exists:
                    a, s = 0.0, 0.0
                    sim.teleport(i, -1000000, -1000000)
                else:
                    tgt = gt[i, t + 1] if targets is None or i not in targets else targets[i][t + 1]
                    a, s = inverse_bicycle(tgt, pos[i], head[i], spd[i], dt)
            sim.set_action(i, a, s)
            rec["accel"][i, t], rec["steer"][i, t] = a, s

    def run_scene(self, scene_idx, scen_json, preproc, logit_steps=(), max_steps=None, replay_only=False):
        """replay_only: no vehicle is policy-controlled - every vehicle is log-replayed through the inverse bicycle model
        (BASELINE config 4; the reference evaluator itself skips scenes without evaluated vehicles)."""
        steps, dt = self.steps, self.dt
        run_steps = steps if max_steps is None else max_steps
        ctx = self.setup_scene(scene_idx, scen_json)
        n, sim, rec = ctx["n"], ctx["sim"], ctx["rec"]
        moving = [i for i in range(n) if ctx["parsed"]["moving"][i]]
        evaluated = random.sample(moving, self.eval_threshold) if len(moving) > self.eval_threshold else moving
        if replay_only:
            evaluated = []
        elif not evaluated:
            return None
        ep = self.new_episode(ctx, evaluated, preproc)
        rec.update(self.policy_record(n))
        rec["evaluated"] = np.array(sorted(evaluated), np.int32)
        next_act = np.zeros((n, 2))
        if self.real_time:
            from . import dense_reward_port as drp
            edges = drp.road_edge_polylines(scen_json)
            rec["dense_reward"] = np.zeros((n, steps + 1, 3))
            tracker = drp.RtgTracker(drp.initial_rtgs(self.w, preproc, n), evaluated, self.max_return, self.min_return)
        for t in range(run_steps):
            self.observe(ctx, t)
            if self.real_time:  # policy_evaluator.py:123-156: RTG of step t from the dense reward of step t-1, then the
                # dense reward of step t; 'nearest_dist' is recorded x max_veh_veh_distance in this mode (evaluator.py:126)
                rec["rtgs"][:, t] = tracker.rtg[0] if t == 0 else tracker.advance(rec["dense_reward"][:, t - 1])
                # QUIRK (evaluator.py:112-113,136-138): the whole reward HISTORY [n, t+1, 8] goes into compute_rewards and
                # all_rewards[i, 0] picks time index 0 - the goal / collision terms of the dense reward are those of STEP 0
                # at every step; only the two distance terms follow the current state
                rec["dense_reward"][:, t], _ = drp.dense_reward_step(self.w, rec["pos"][:, t], rec["existence"][:, t],
                                                                     rec["reward"][:, 0], edges)
                rec["nearest_dist"][:, t] *= self.w.max_veh_veh_distance
                rec["gt_nearest_dist"][:, t] *= self.w.max_veh_veh_distance
            self.update_state(ep, ctx, rec, t)
            self.predict_step(ep, rec, t, scene_idx, next_act, logit_steps)
            self.apply_controls(ctx, t, evaluated, next_act)
            sim.step(dt)
        if run_steps == steps:
            self.observe(ctx, steps)
            if self.real_time:
                rec["dense_reward"][:, steps], _ = drp.dense_reward_step(self.w, rec["pos"][:, steps], rec["existence"][:, steps],
                                                                         rec["reward"][:, 0], edges)
                rec["nearest_dist"][:, steps] *= self.w.max_veh_veh_distance
                rec["gt_nearest_dist"][:, steps] *= self.w.max_veh_veh_distance
            if evaluated:
                self.metrics.add_scene(rec, evaluated)
        return rec
