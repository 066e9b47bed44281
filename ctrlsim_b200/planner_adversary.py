"""Planner-vs-adversary evaluation on the batched hot path.

Mirror of ``PlannerAdversaryEvaluator`` (evaluators/planner_adversary_evaluator.py:26-593) and of its driver
eval_planner.py: in every scene ONE vehicle (the ego / SDC) is driven by the *planner* policy and ONE by the
*adversary* - either a second CtRL-Sim policy with its own tilts (cfgs/policy/ctrl_sim_adversary.yaml: veh_veh_tilt
-10) and possibly its own checkpoint, or the scripted CAT trajectory (``adversary.name == 'cat'``,
planner_adversary_evaluator.py:165-198) pushed through the inverse bicycle model; every other vehicle is log-replayed.

Nothing new runs on the device: both policies are ``B200Policy`` objects stepping the same C-ABI calls as the
single-policy evaluator, each on its own ``SceneBatch.policy_view`` (own RTG history, context membership, focal
groups, sampled bins - what the reference keeps per Policy object and per ``key_dict``) of one shared world (simulator
state, state/action history, trace).  A third view carries the merged controls into ``ctrlsim_sim_step``.  Scenes
are batched and sharded over ranks like in ``B200PolicyEvaluator``; the per-scene statistics are tiny and are
exchanged with one ``all_gather_object``.
"""
from __future__ import annotations

import os
import pickle
import json

import numpy as np
import torch

from . import lib as _lib
from .batch import SceneBatch
from .evaluator import B200Policy
from .scenario import parse_scenario


class CatAdversary:
    """Stand-in for the reference's policy object of cfgs/policy/cat.yaml: no network, the adversary follows the
    pre-computed trajectory ``adv_traj`` of its scene (planner_adversary_evaluator.py:165-198,519-523)."""
    name = "cat"
    real_time_rewards = False
    model_path = "cat"


def moving_average(data, window_size):
    """utils/sim.py:198-202"""
    interval = np.pad(data, window_size // 2, "edge")
    return np.convolve(interval, np.ones(int(window_size)) / float(window_size), "valid")


def adversary_trajectory(adv_pos):
    """adv_pos [T,2] -> [T,5] x, y, vx, vy, yaw (get_polyline_yaw / get_polyline_vel, utils/sim.py:204-222;
    planner_adversary_evaluator.py:455-458)."""
    adv_pos = np.asarray(adv_pos, np.float64)
    post = np.roll(adv_pos, shift=-1, axis=0)
    diff = post - adv_pos
    yaw = np.arctan2(diff[:, 1], diff[:, 0])
    yaw[-1] = yaw[-2]
    for i in range(len(yaw) - 1):
        if yaw[i + 1] - yaw[i] > 1.5 * np.pi:
            yaw[i + 1] -= 2 * np.pi
        elif yaw[i] - yaw[i + 1] > 1.5 * np.pi:
            yaw[i + 1] += 2 * np.pi
    yaw = moving_average(yaw, 5)
    post_v = np.roll(adv_pos, shift=-1, axis=0)
    post_v[-1] = adv_pos[-1]
    vel = (post_v - adv_pos) / 0.1
    return np.concatenate([adv_pos, vel, yaw.reshape(-1, 1)], axis=1)


def vehicle_index_of_object(scen_json, obj_idx):
    """Index, among the simulated vehicles, of JSON object ``obj_idx`` (the reference matches the object's first
    position against the simulator's vehicles, planner_adversary_evaluator.py:431-441): vehicles are the objects of
    type 'vehicle' that are valid at step 0, in file order.  None if that object is not simulated."""
    k = -1
    for i, o in enumerate(scen_json["objects"]):
        if bool(o["valid"][0]) and o["type"] == "vehicle":
            k += 1
            if i == obj_idx:
                return k
        elif i == obj_idx:
            return None
    return None


class PlannerAdversaryStats:
    """update_running_statistics / compute_metrics of the reference (planner_adversary_evaluator.py:201-372,374-428)
    on per-scene arrays; host numpy (two vehicles per scene)."""
    LIST_KEYS = ("ades", "fdes", "goal", "progress", "cr", "cr_w_adv", "offroad", "jerk", "steer_rate", "accel",
                 "lin_sim", "lin_gt", "ang_sim", "ang_gt", "acc_sim", "acc_gt", "near_sim", "near_gt", "coll_speed")

    def __init__(self, cfg):
        self.cfg = cfg
        self.w = cfg.dataset.waymo
        self.steps, self.dt = cfg.nocturne.steps, cfg.nocturne.dt
        self.hist = cfg.eval_planner_adversary.history_steps
        self.data = {k: [] for k in self.LIST_KEYS}

    def add_scene(self, rec, ego, adv):
        """rec: arrays indexed [vehicle, t]: pos [n,T1,2], vel, heading, existence, accel, steer, reward [n,T1,8],
        nearest_dist, gt_nearest_dist, gt_pos, gt_heading, gt_speed, gt_accel, size [n,2]."""
        d, T1, hist, dt = self.data, self.steps + 1, self.hist, self.dt
        future = np.zeros(T1, bool)
        future[hist:] = True
        ego_mask = rec["existence"][ego].astype(bool) & future
        adv_mask = rec["existence"][adv].astype(bool) & future
        coll_sc, coll_adv_sc, off_sc = [], [], []
        hit = False
        if ego_mask.sum() != 0:
            rew = rec["reward"][ego][ego_mask]
            goal_achieved = bool(np.any(np.sum(rew[:, :1], axis=1) == 1))
            d["goal"].append(float(goal_achieved))
            coll_sc.append(float(np.any(rew[:, 6] == 1)))
            off_sc.append(float(np.any(rew[:, 7] == 1)))
            sim, gt = rec["pos"][ego].astype(np.float64), rec["gt_pos"][ego].astype(np.float64)
            d["ades"].append(np.linalg.norm(sim[ego_mask] - gt[ego_mask], axis=1).mean())
            last = int(np.where(ego_mask)[0][-1])
            d["fdes"].append(np.linalg.norm(sim[last] - gt[last]))
            seg = sim[hist:last + 1]
            if goal_achieved:
                progress = np.linalg.norm(np.diff(seg, axis=0), axis=-1).sum()
            else:
                to_goal = np.linalg.norm(seg - gt[last][None], axis=-1)
                closer = np.diff(to_goal) < 0
                progress = np.linalg.norm(np.diff(seg, axis=0), axis=-1)[closer].sum()
            d["progress"].append(progress)
            acc = rec["accel"][ego][ego_mask]
            d["jerk"].append(np.abs(np.diff(acc)) / dt)
            d["accel"].append(np.abs(acc))
            d["steer_rate"].append(np.abs(np.diff(rec["steer"][ego][ego_mask])) / dt)
        if adv_mask.sum() != 0:
            v = rec["vel"][adv].astype(np.float64)[adv_mask]
            d["lin_sim"].append(np.linalg.norm(v, axis=1))
            d["lin_gt"].append(rec["gt_speed"][adv][adv_mask])
            d["ang_sim"].append(rec["heading"][adv][adv_mask] / dt)
            d["ang_gt"].append(rec["gt_heading"][adv][adv_mask] / dt)
            ga, sa = rec["gt_accel"][adv][adv_mask], rec["accel"][adv][adv_mask]
            keep = np.ones(ga.shape, bool)
            keep[0] = keep[-1] = False
            d["acc_gt"].append(ga[keep])
            d["acc_sim"].append(sa[keep])
            d["near_gt"].append(rec["gt_nearest_dist"][adv][adv_mask])
            d["near_sim"].append(rec["nearest_dist"][adv][adv_mask])
        if ego_mask.sum() != 0 and adv_mask.sum() != 0:
            ec, ac = rec["reward"][ego][ego_mask, 6], rec["reward"][adv][adv_mask, 6]
            m = min(len(ec), len(ac))
            ec, ac = ec[:m], ac[:m]
            both = ((ec == ac).astype(float) * ec).astype(bool)
            has = float(np.any(both))
            if has == 1.0:
                ep = rec["pos"][ego].astype(np.float64)[ego_mask][:m]
                ap = rec["pos"][adv].astype(np.float64)[adv_mask][:m]
                avx, avy = rec["vel"][adv][:, 0].astype(np.float64), rec["vel"][adv][:, 1].astype(np.float64)
                valid = False
                for c in np.where(both)[0]:
                    if np.linalg.norm(ep[c] - ap[c]) < rec["size"][ego, 0] + rec["size"][adv, 0]:
                        valid = True
                        # quirk kept: x is indexed in the UNMASKED series, y is not indexed at all (:352)
                        d["coll_speed"].append(np.sqrt(avx[c] ** 2 + avy ** 2))
                        break
                if not valid:
                    has = 0.0
            coll_adv_sc.append(has)
            hit = bool(has)
        if coll_sc:
            d["cr"].append(np.mean(coll_sc))
            d["cr_w_adv"].append(np.mean(coll_adv_sc if coll_adv_sc else [0.0]))
            d["offroad"].append(np.mean(off_sc))
        return hit

    def merge(self, others):
        for o in others:
            for k in self.LIST_KEYS:
                self.data[k].extend(o[k])

    @staticmethod
    def _jsd(p, q):
        """scipy.spatial.distance.jensenshannon (natural log) on two already normalised histograms"""
        p, q = np.asarray(p, np.float64), np.asarray(q, np.float64)
        p, q = p / p.sum(), q / q.sum()
        m = (p + q) / 2.0
        left = np.where(p > 0, p * np.log(np.where(p > 0, p, 1.0) / np.where(m > 0, m, 1.0)), 0.0)
        right = np.where(q > 0, q * np.log(np.where(q > 0, q, 1.0) / np.where(m > 0, m, 1.0)), 0.0)
        return float(np.sqrt((left.sum() + right.sum()) / 2.0))

    def compute(self):
        d, w = self.data, self.w
        mean = lambda xs: float(np.array(xs).mean()) if len(xs) else float("nan")
        cat = lambda xs: np.concatenate([np.asarray(x, np.float64).reshape(-1) for x in xs]) if len(xs) else np.zeros(0)
        m = {"ego_goal": mean(d["goal"]), "ego_prog": mean(d["progress"]), "ego_cr": mean(d["cr"]),
             "ego_cr_w_adv": mean(d["cr_w_adv"]), "ego_or": mean(d["offroad"]), "ego_fde": mean(d["fdes"]),
             "ego_ade": mean(d["ades"]), "ego_accel": mean(cat(d["accel"])), "ego_jerk": mean(cat(d["jerk"])),
             "ego_steer_rate": mean(cat(d["steer_rate"])), "adv_coll_speed": mean(cat(d["coll_speed"]))}

        def hist_jsd(sim, gt, edges):
            sim, gt = cat(sim), cat(gt)
            return self._jsd(np.histogram(sim, bins=edges)[0] / len(sim), np.histogram(gt, bins=edges)[0] / len(gt))

        clip = lambda xs, lo, hi: [np.clip(x, lo, hi) for x in xs]
        m["adv_lin_jsd"] = hist_jsd(clip(d["lin_sim"], 0, 30), clip(d["lin_gt"], 0, 30), np.arange(201) * 0.5 * (100 / 30))
        m["adv_ang_jsd"] = hist_jsd(clip(d["ang_sim"], -50, 50), clip(d["ang_gt"], -50, 50), np.arange(201) * 0.5 - 50)
        gt_acc = cat(d["acc_gt"])
        gt_acc = (np.clip(gt_acc, w.min_accel, w.max_accel) - w.min_accel) / (w.max_accel - w.min_accel)
        gt_acc = np.round(gt_acc * (w.accel_discretization - 1)) / (w.accel_discretization - 1)
        gt_acc = gt_acc * (w.max_accel - w.min_accel) + w.min_accel
        m["adv_acc_jsd"] = hist_jsd(d["acc_sim"], [gt_acc], np.arange(w.accel_discretization + 1) * 2 - w.accel_discretization)
        m["nearest_dist_jsd"] = hist_jsd(clip(d["near_sim"], 0, 40), clip(d["near_gt"], 0, 40), np.arange(201) * 0.5 * (100 / 40))
        return m


class B200PlannerAdversaryEvaluator:
    def __init__(self, cfg, planner: B200Policy, adversary, scenes=None, pairs=None, adv_trajs=None, scene_ids=None):
        """``planner``: B200Policy; ``adversary``: B200Policy or CatAdversary.  Give the two policies separate
        ``DeviceModel`` objects (the reference loads two checkpoints, eval_planner.py:57,120): the step caches live
        in the model handle and are keyed by consecutive steps of one policy - two policies sharing a handle stay
        exact but recompute every step from scratch.
        ``scenes`` / ``pairs`` / ``adv_trajs`` (optional, in memory): per scene the {'json','preproc'} dict, the
        (ego, adversary) vehicle indices and - for the CAT adversary - the adversary's positions [steps+1, 2].
        Otherwise the reference's files are read: ``cfg.cat.dict_path`` (eval_planner_dict.pkl),
        ``cfg.nocturne_waymo_val_interactive_folder`` and <dataset_root>/preprocess/val_interactive
        (planner_adversary_evaluator.py:31-41,430-460,476-490)."""
        self.cfg, self.planner, self.adversary = cfg, planner, adversary
        self.steps, self.dt = cfg.nocturne.steps, cfg.nocturne.dt
        self.history_steps = cfg.eval_planner_adversary.history_steps
        self.rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        self.world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        self.is_cat = getattr(adversary, "name", "") == "cat"
        if not self.is_cat and not isinstance(adversary, B200Policy):
            raise TypeError("adversary must be a B200Policy or a CatAdversary")
        if not self.is_cat and adversary.seed == planner.seed:
            # the explicit sampler is keyed by (seed; scene, vehicle, step, component): with equal seeds the two policies
            # would draw IDENTICAL Philox bits for the RTG of the same (scene, vehicle, step) - correlated draws where
            # the reference's two policies consume independent stretches of the torch generator
            raise ValueError("planner and adversary need distinct sampler seeds (B200Policy(..., seed=...))")
        if scenes is None:
            scenes, pairs, adv_trajs, scene_ids = self._load_files()
        if pairs is None or len(pairs) != len(scenes):
            raise ValueError("one (ego, adversary) pair per scene is required")
        if self.is_cat and (adv_trajs is None or any(a is None for a in adv_trajs)):
            raise ValueError("the CAT adversary needs an adversary trajectory for every scene")
        ids = list(range(len(scenes))) if scene_ids is None else list(scene_ids)
        keep = [k for k in range(len(scenes)) if k % self.world == self.rank]
        self.scenes = [scenes[k] for k in keep]
        self.pairs = [tuple(int(x) for x in pairs[k]) for k in keep]
        self.adv_trajs = None if adv_trajs is None else [adv_trajs[k] for k in keep]
        self.scene_ids = [ids[k] for k in keep]
        self.batch = self.view_planner = self.view_adversary = self.view_control = None

    def _load_files(self):
        cfg = self.cfg
        with open(cfg.cat.dict_path, "rb") as f:
            table = pickle.load(f)
        scenes, pairs, trajs, ids = [], [], [], []
        for i, k in enumerate(table):
            if len(scenes) == cfg.eval_planner_adversary.num_files_to_evaluate:
                break
            ent = table[k]
            name = ent["nocturne_path"][66:]  # the reference strips its own absolute prefix (:37)
            pkl = os.path.join(cfg.dataset_root, "preprocess/val_interactive", f"{name[:-5]}_physics.pkl")
            if not os.path.exists(pkl) or "adv_traj" not in ent:
                continue  # skipped exactly like the reference (:486-487, :447-448)
            with open(os.path.join(cfg.nocturne_waymo_val_interactive_folder, name)) as f:
                js = json.load(f)
            ego = vehicle_index_of_object(js, ent["nocturne_sdc_id"])
            adv = vehicle_index_of_object(js, ent["nocturne_adversary_id"])
            if ego is None or adv is None:
                continue
            with open(pkl, "rb") as f:
                pre = pickle.load(f)
            scenes.append({"name": name, "json": js, "preproc": pre})
            pairs.append((ego, adv))
            trajs.append(np.asarray(ent["adv_traj"], np.float64))
            ids.append(i)
        return scenes, pairs, trajs, ids

    # ------------------------------------------------------------------------------------------------------------------
    def build_batch(self):
        cfg = self.cfg
        sc = cfg.nocturne["scenario"]
        parsed = [parse_scenario(s["json"], self.steps, sc["moving_threshold"], sc["speed_threshold"]) for s in self.scenes]
        dev = self.planner.model.device
        both = [[e, a] for e, a in self.pairs]
        self.batch = SceneBatch(cfg, self.scenes, self.scene_ids, dev, eval_threshold=2, parsed=parsed, evaluated_sets=both)
        self.view_planner = self.batch.policy_view([[e] for e, _ in self.pairs])
        gt = None
        if self.is_cat:
            # apply_adv_traj == apply_gt_action with the target state taken from adv_traj (:165-198): give the control
            # view a copy of the log-replay targets whose adversary rows hold the scripted states from step
            # history_steps on (before that the adversary is log-replayed like everyone else, :515-526)
            gt = self.batch.t["gt"].cpu().numpy().copy()
            for s, ((_, adv), pos) in enumerate(zip(self.pairs, self.adv_trajs)):
                tr = adversary_trajectory(pos)
                T1 = min(self.steps + 1, len(tr))
                h = self.history_steps
                gt[s, adv, h:T1, 0], gt[s, adv, h:T1, 1] = tr[h:T1, 0], tr[h:T1, 1]
                gt[s, adv, h:T1, 2] = tr[h:T1, 4]
                gt[s, adv, h:T1, 3] = np.sqrt(tr[h:T1, 2] ** 2 + tr[h:T1, 3] ** 2)
            self.view_adversary = None
            self.view_control = self.batch.policy_view([[e] for e, _ in self.pairs], gt=gt)
        else:
            self.view_adversary = self.batch.policy_view([[a] for _, a in self.pairs])
            self.view_control = self.batch.policy_view(both)
        return self.batch

    def rollout(self, max_steps=None):
        """The step loop of evaluate_planner_adversary (:505-540), all scenes at once."""
        if self.batch is None:
            self.build_batch()
        P, A = self.planner, self.adversary
        vp, va, vc = self.view_planner, self.view_adversary, self.view_control
        P.reset(self.batch)  # world + the shared history; then the per-policy state of every view
        for v in (vp, va, vc):
            if v is not None:
                v.reset_dynamic()
        if va is not None and A.model is not P.model:
            A.attach_caches(self.batch)
        steps = self.steps if max_steps is None else max_steps
        st = torch.cuda.current_stream(P.model.device).cuda_stream
        ego_mask = vp.t["evaluated"].bool().unsqueeze(-1)
        for t in range(steps):
            P.update_state(vp, t)  # update_vehicle_data_dict + both update_state calls: the world record is shared
            P.predict(vp, t)
            if va is not None:
                A.predict(va, t)
                torch.where(ego_mask, vp.t["next_action"], va.t["next_action"], out=vc.t["next_action"])
            else:
                vc.t["next_action"].copy_(vp.t["next_action"])
            _lib.check(P.lib.ctrlsim_sim_step(P.model.handle, vc.ptr, t, st), "ctrlsim_sim_step")
        if steps == self.steps:
            P.update_state(vp, self.steps)
        return self.batch

    def records(self):
        """Per-scene host arrays in the layout PlannerAdversaryStats.add_scene reads."""
        b, dt, T = self.batch, self.dt, self.steps
        tr = b.trace()
        gt = b.t["gt"].cpu().numpy().astype(np.float64)
        size = np.stack([b.t["veh_len"].cpu().numpy(), b.t["veh_wid"].cpu().numpy()], -1).astype(np.float64)
        out = []
        for s in range(b.S):
            n = int(tr["n_veh"][s])
            gt_acc = np.zeros((n, T + 1))
            gt_acc[:, 1:T - 1] = (gt[s, :n, 2:T, 3] - gt[s, :n, 0:T - 2, 3]) / (2 * dt)
            out.append({"pos": tr["tr_pos"][s, :n], "vel": tr["tr_vel"][s, :n],
                        "heading": tr["tr_heading"][s, :n].astype(np.float64),
                        "existence": tr["tr_exist"][s, :n].astype(np.float64),
                        "accel": tr["tr_action"][s, :n, :, 0], "steer": tr["tr_action"][s, :n, :, 1],
                        "reward": tr["tr_reward"][s, :n].astype(np.float64),
                        "nearest_dist": tr["tr_nearest"][s, :n, :, 0], "gt_nearest_dist": tr["tr_nearest"][s, :n, :, 1],
                        "gt_pos": gt[s, :n, :, :2], "gt_heading": gt[s, :n, :, 2], "gt_speed": gt[s, :n, :, 3],
                        "gt_accel": gt_acc, "size": size[s, :n]})
        return out

    def gather(self, stats: PlannerAdversaryStats) -> PlannerAdversaryStats:
        """The one exchange of an evaluation: every rank contributes the per-scene lists of its scenes (a few hundred
        floats per scene) and all ranks end up with the statistics of the whole job.  Every metric is a mean or a
        histogram over these lists, so the rank order of the concatenation only matters at float-summation level."""
        if self.world == 1:
            return stats
        parts = [None] * self.world
        torch.distributed.all_gather_object(parts, stats.data)
        merged = PlannerAdversaryStats(self.cfg)
        merged.merge(parts)
        return merged

    def evaluate_planner_adversary(self):
        self.rollout()
        stats = PlannerAdversaryStats(self.cfg)
        for rec, (ego, adv) in zip(self.records(), self.pairs):
            stats.add_scene(rec, ego, adv)
        m = self.gather(stats).compute()
        return m, ["{}: {:.6f}".format(k, v) for k, v in m.items()]
