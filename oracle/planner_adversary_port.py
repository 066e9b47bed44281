"""ORACLE (test infrastructure only): numpy restatement of the reference's planner-vs-adversary evaluation of one scene
(SURVEY 8(f) N2).  Never imported by the product path.

Restated here (reference file:line):
  loop      evaluate_planner_adversary                evaluators/planner_adversary_evaluator.py:466-593
  CAT       apply_adv_traj, get_planner_adversary     evaluators/planner_adversary_evaluator.py:165-198,430-460
            get_polyline_yaw / get_polyline_vel       utils/sim.py:198-222
  metrics   update_running_statistics, compute_metrics evaluators/planner_adversary_evaluator.py:201-428
Two ``RolloutPort`` objects (oracle/policy_port.py) play the two Policy objects: each has its own buffers, relevant
agent sets and RTG series (the reference's ``key_dict``: 'planner_rtgs' / 'adversary_rtgs') and serves exactly one
vehicle; the world record (vehicle_data_dict) and the simulator are shared.  Pinned against the reference itself by
tests/golden/planner_adversary_*.npz (oracle/make_golden_planner_adversary.py).
"""
from __future__ import annotations

import numpy as np

from . import policy_port


def moving_average(data, window_size):  # utils/sim.py:198-202
    interval = np.pad(data, window_size // 2, "edge")
    window = np.ones(int(window_size)) / float(window_size)
    return np.convolve(interval, window, "valid")


def polyline_yaw(polyline):  # utils/sim.py:204-215
    post = np.roll(polyline, shift=-1, axis=0)
    diff = post - polyline
    yaw = np.arctan2(diff[:, 1], diff[:, 0])
    yaw[-1] = yaw[-2]
    for i in range(len(yaw) - 1):
        if yaw[i + 1] - yaw[i] > 1.5 * np.pi:
            yaw[i + 1] -= 2 * np.pi
        elif yaw[i] - yaw[i + 1] > 1.5 * np.pi:
            yaw[i + 1] += 2 * np.pi
    return moving_average(yaw, window_size=5)


def polyline_vel(polyline):  # utils/sim.py:217-222
    post = np.roll(polyline, shift=-1, axis=0)
    post[-1] = polyline[-1]
    return (post - polyline) / 0.1


class PlannerAdversaryMetricsPort:
    """Lists exactly as the reference keeps them (planner_adversary_evaluator.py:45-75)."""

    def __init__(self, cfg, history_steps=10):
        self.cfg, self.w = cfg, cfg.dataset.waymo
        self.steps, self.dt, self.hist = cfg.nocturne.steps, cfg.nocturne.dt, history_steps
        for k in ("ades", "fdes", "goal", "progress", "cr", "cr_adv", "off", "jerk", "steer_rate", "accel", "lin_sim",
                  "lin_gt", "ang_sim", "ang_gt", "acc_sim", "acc_gt", "near_sim", "near_gt", "coll_speed"):
            setattr(self, k, [])

    def add_scene(self, rec, ego, adv):
        T1, hist, dt = self.steps + 1, self.hist, self.dt
        future = np.zeros(T1).astype(bool)
        future[hist:] = True
        coll, coll_adv, off = [], [], []
        ego_mask = rec["existence"][ego].astype(bool) * future
        if ego_mask.sum() != 0:
            rew = rec["reward"][ego][ego_mask]
            goal_achieved = np.any(np.sum(rew[:, :1], axis=1) == 1)
            self.goal.append(float(goal_achieved))
            coll.append(float(np.any(rew[:, 6] == 1)))
            off.append(float(np.any(rew[:, 7] == 1)))
            sim, gt = rec["pos"][ego], rec["gt_pos"][ego]
            self.ades.append(np.linalg.norm(sim[ego_mask] - gt[ego_mask], axis=1).mean())
            last = np.where(ego_mask == 1)[-1][-1]
            self.fdes.append(np.linalg.norm(sim[last] - gt[last]))
            if goal_achieved:
                progress = np.linalg.norm(np.diff(sim[hist:last + 1], axis=0), axis=-1).sum()
            else:
                dist_to_goal = np.linalg.norm(sim[hist:last + 1] - np.expand_dims(gt[last], axis=0), axis=-1)
                closer = np.diff(dist_to_goal) < 0
                progress = np.linalg.norm(np.diff(sim[hist:last + 1], axis=0), axis=-1)[closer].sum()
            self.progress.append(progress)
            acc = rec["accel"][ego][ego_mask]
            self.jerk.append(np.abs(np.diff(acc)) / dt)
            self.accel.append(np.abs(acc))
            self.steer_rate.append(np.abs(np.diff(rec["steer"][ego][ego_mask])) / dt)
        adv_mask = rec["existence"][adv].astype(bool) * future
        if adv_mask.sum() != 0:
            self.lin_sim.append(np.linalg.norm(rec["vel"][adv][adv_mask], axis=1)[:, None])
            self.lin_gt.append(rec["gt_speed"][adv][adv_mask][:, None])
            self.ang_sim.append((rec["heading"][adv][adv_mask] / dt)[:, None])
            self.ang_gt.append((rec["gt_heading"][adv][adv_mask] / dt)[:, None])
            ga, sa = rec["gt_accel"][adv][adv_mask], rec["accel"][adv][adv_mask]
            m = np.ones(ga.shape).astype(bool)
            m[0] = False
            m[-1] = False
            self.acc_sim.append(sa[m][:, None])
            self.acc_gt.append(ga[m][:, None])
            self.near_gt.append(rec["gt_nearest_dist"][adv][adv_mask][:, None])
            self.near_sim.append(rec["nearest_dist"][adv][adv_mask][:, None])
        if ego_mask.sum() != 0 and adv_mask.sum() != 0:
            ec, ac = rec["reward"][ego][ego_mask, 6], rec["reward"][adv][adv_mask, 6]
            k = min(len(ec), len(ac))
            ec, ac = ec[:k], ac[:k]
            both = ((ec == ac).astype(float) * ec).astype(bool)
            has = float(np.any(both))
            if has == 1.0:
                ego_pos = rec["pos"][ego][ego_mask][:k]
                adv_pos = rec["pos"][adv][adv_mask][:k]
                avx, avy = rec["vel"][adv][:, 0], rec["vel"][adv][:, 1]
                ok = False
                for c in np.where(both)[0]:
                    if np.linalg.norm(ego_pos[c] - adv_pos[c]) < rec["size"][ego, 0] + rec["size"][adv, 0]:
                        ok = True
                        self.coll_speed.append(np.sqrt(avx[c] ** 2 + avy ** 2))  # sic (:352)
                        break
                if not ok:
                    has = 0.0
            coll_adv.append(has)
        if len(coll) > 0:
            self.cr.append(np.array(coll).mean())
            if len(coll_adv) == 0:
                coll_adv.append(0.0)
            self.cr_adv.append(np.array(coll_adv).mean())
            self.off.append(np.array(off).mean())

    def compute(self):
        w = self.w
        jsd = policy_port.MetricsPort.jsd
        with np.errstate(invalid="ignore"):
            m = {"ego_goal": np.array(self.goal).mean(), "ego_prog": np.array(self.progress).mean(),
                 "ego_cr": np.array(self.cr).mean(), "ego_cr_w_adv": np.array(self.cr_adv).mean(),
                 "ego_or": np.array(self.off).mean(), "ego_fde": np.array(self.fdes).mean(),
                 "ego_ade": np.array(self.ades).mean(), "ego_accel": np.concatenate(self.accel, axis=0).mean(),
                 "ego_jerk": np.concatenate(self.jerk, axis=0).mean(),
                 "ego_steer_rate": np.concatenate(self.steer_rate, axis=0).mean(),
                 "adv_coll_speed": np.array(self.coll_speed).mean() if len(self.coll_speed) else float("nan")}

        def hj(sim, gt, edges):
            return jsd(np.histogram(sim, bins=edges)[0] / len(sim), np.histogram(gt, bins=edges)[0] / len(gt))

        lg, ls = np.clip(np.concatenate(self.lin_gt, 0), 0, 30), np.clip(np.concatenate(self.lin_sim, 0), 0, 30)
        m["adv_lin_jsd"] = hj(ls, lg, np.arange(201) * 0.5 * (100 / 30))
        ag, as_ = np.clip(np.concatenate(self.ang_gt, 0), -50, 50), np.clip(np.concatenate(self.ang_sim, 0), -50, 50)
        m["adv_ang_jsd"] = hj(as_, ag, np.arange(201) * 0.5 - 50)
        acc_gt = np.concatenate(self.acc_gt, 0)
        acc_gt = (np.clip(acc_gt, a_min=w.min_accel, a_max=w.max_accel) - w.min_accel) / (w.max_accel - w.min_accel)
        acc_gt = np.round(acc_gt * (w.accel_discretization - 1))
        acc_gt /= (w.accel_discretization - 1)
        acc_gt = (acc_gt * (w.max_accel - w.min_accel)) + w.min_accel
        m["adv_acc_jsd"] = hj(np.concatenate(self.acc_sim, 0), acc_gt,
                              np.arange(w.accel_discretization + 1) * 2 - w.accel_discretization)
        ng, ns = np.clip(np.concatenate(self.near_gt, 0), 0, 40), np.clip(np.concatenate(self.near_sim, 0), 0, 40)
        m["nearest_dist_jsd"] = hj(ns, ng, np.arange(201) * 0.5 * (100 / 40))
        return {k: float(v) for k, v in m.items()}


class PlannerAdversaryPort:
    def __init__(self, cfg, planner: policy_port.RolloutPort, adversary, history_steps=10):
        """``adversary``: a RolloutPort, or None for the scripted CAT adversary."""
        self.cfg, self.planner, self.adversary = cfg, planner, adversary
        self.steps, self.dt, self.hist = cfg.nocturne.steps, cfg.nocturne.dt, history_steps
        self.metrics = PlannerAdversaryMetricsPort(cfg, history_steps)

    def run_scene(self, scene_idx, scen_json, preproc, ego, adv, adv_pos=None, max_steps=None):
        P, A = self.planner, self.adversary
        steps = self.steps
        run_steps = steps if max_steps is None else max_steps
        ctx = P.setup_scene(scene_idx, scen_json)
        n, sim, rec, gt = ctx["n"], ctx["sim"], ctx["rec"], ctx["gt"]
        ep_p, rec_p = P.new_episode(ctx, [ego], preproc), P.policy_record(n)
        targets = None
        if A is not None:
            ep_a, rec_a = A.new_episode(ctx, [adv], preproc), A.policy_record(n)
            controlled = [ego, adv]
        else:
            pos = np.asarray(adv_pos, np.float64)
            traj = np.concatenate((pos, polyline_vel(pos), polyline_yaw(pos).reshape(-1, 1)), axis=1)
            tgt = gt[adv].copy()  # rows (x, y, heading, speed, exist, .., length): BicycleModel(x, y, theta, vel, L)
            k = min(len(traj), steps + 1)
            tgt[:k, 0], tgt[:k, 1], tgt[:k, 2] = traj[:k, 0], traj[:k, 1], traj[:k, 4]
            tgt[:k, 3] = np.sqrt(traj[:k, 2] ** 2 + traj[:k, 3] ** 2)
            rec_a = None
            controlled = [ego]
        next_act = np.zeros((n, 2))  # planner and adversary write disjoint rows (their own key_dict slots)
        for t in range(run_steps):
            P.observe(ctx, t)
            P.update_state(ep_p, ctx, rec_p, t)
            if A is not None:
                A.update_state(ep_a, ctx, rec_a, t)
            P.predict_step(ep_p, rec_p, t, scene_idx, next_act)
            if A is not None:
                A.predict_step(ep_a, rec_a, t, scene_idx, next_act)
                tg = None
            else:
                # the scripted trajectory only takes over from step history_steps - 1 (:515-526)
                tg = {adv: tgt} if t >= self.hist - 1 else None
            P.apply_controls(ctx, t, controlled, next_act, targets=tg)
            sim.step(self.dt)
        if run_steps == steps:
            P.observe(ctx, steps)
            self.metrics.add_scene(rec, ego, adv)
        rec["planner"], rec["adversary"] = rec_p, rec_a
        rec["ego"], rec["adv"] = ego, adv
        return rec
