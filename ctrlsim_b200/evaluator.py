"""Drop-in policy / evaluator pair for the closed-loop rollout, batched over scenes on one B200.

Mirrors the reference's plugin API for this path:

  B200Policy          <-> policies.AutoregressivePolicy   (policies/policy.py:8-154, autoregressive_policy.py:10-274)
      same constructor arguments, same attribute names the evaluator reads, and the same four verbs
      reset / update_state / predict / act - operating on a whole SceneBatch instead of one vehicle_data_dict.
  B200PolicyEvaluator <-> evaluators.PolicyEvaluator      (evaluators/policy_evaluator.py:27-44,426-595)
      ``B200PolicyEvaluator(cfg, policy).evaluate_policy() -> (metrics_dict, [str])`` with the reference's keys:
      goal, collision_rate, offroad_rate, fde, ade, lin_speed_jsd, ang_speed_jsd, accel_jsd, nearest_dist_jsd.

Host code is Python; every per-step operation is a CUDA kernel behind the C-ABI (ctrlsim_b200/lib.py).  Scenes shard
over ranks (one process per GPU); the only collective is one all-reduce of the summary buffer per evaluation.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import pickle
import random
import warnings

import numpy as np
import torch

from . import lib as _lib
from .batch import SceneBatch
from .model import DeviceModel


class B200Policy:
    def __init__(self, cfg, model_path, model, use_rtg=True, predict_rtgs=True, discretize_rtgs=True,
                 real_time_rewards=False, privileged_return=False, max_return=False, min_return=False, key_dict=None,
                 tilt_dict=None, name="ctrl_sim", action_temperature=1.0, nucleus_sampling=False,
                 nucleus_threshold=0.8, seed=0, chunk_groups=128):
        if not isinstance(model, DeviceModel):
            raise TypeError("B200Policy needs a ctrlsim_b200.model.DeviceModel (weights resident on the GPU)")
        # Supported switch combinations (cfgs/policy/*.yaml, policies/policy.py:9-39):
        #   ctrl_sim  use_rtg, predict_rtgs, discretize_rtgs                      RTG head + sampling, two passes
        #   dt        use_rtg, real_time_rewards, not predict_rtgs, not discretize_rtgs (+ max_return / min_return) with
        #             the decision-transformer network (cfgs/model/dt.yaml): RTGs tracked from the dense reward
        #   the same tracked RTGs, discretised, with the CtRL-Sim network (predict_rtgs=False, discretize_rtgs=True)
        dt_model = bool(cfg.model.get("decision_transformer", False))
        if not use_rtg and not (dt_model and real_time_rewards and not predict_rtgs):
            # use_rtg=False with the decision-transformer network is what cfgs/policy/dt.yaml composes to AS SHIPPED (it
            # spells the key `use_rtgs`, so cfgs/policy/base.yaml's use_rtg: False stays): the tracked RTGs never reach
            # the policy's buffers (policies/policy.py:89-95) and the network is fed RTG (0, 0, 0) - supported
            raise NotImplementedError("use_rtg=False is implemented for the decision-transformer baseline only (the il / "
                                      "trajeglish baselines use other token layouts)")
        if predict_rtgs and (real_time_rewards or max_return or min_return or not discretize_rtgs or dt_model):
            raise ValueError("predict_rtgs=True needs the CtRL-Sim network with discretize_rtgs=True and no real_time_rewards "
                             "/ max_return / min_return (cfgs/policy/ctrl_sim.yaml)")
        if not predict_rtgs and not real_time_rewards:
            raise NotImplementedError("RTGs that are neither predicted nor tracked in real time: the reference then feeds "
                                      "zeros (policies/policy.py:89-95); not implemented")
        if real_time_rewards and discretize_rtgs == dt_model:
            raise ValueError("discretize_rtgs must be False for the decision-transformer network (continuous RTG inputs, "
                             "cfgs/policy/dt.yaml) and True for the CtRL-Sim network (RTG embedding tables)")
        self.cfg = cfg.copy()
        self.model_path, self.model, self.name = model_path, model, name
        self.model.eval()
        self.cfg_model, self.cfg_rl_waymo = cfg.model, cfg.dataset.waymo
        self.steps = self.cfg.nocturne.steps
        self.use_rtg, self.predict_rtgs, self.discretize_rtgs = use_rtg, predict_rtgs, discretize_rtgs
        self.real_time_rewards, self.privileged_return = real_time_rewards, privileged_return
        self.max_return, self.min_return = max_return, min_return
        self.key_dict = key_dict or {"next_acceleration": "next_acceleration", "next_steering": "next_steering",
                                     "rtgs": "rtgs"}
        self.tilt_dict = tilt_dict or {"tilt": True, "goal_tilt": 0, "veh_veh_tilt": 0, "veh_edge_tilt": 0}
        self.action_temperature = action_temperature
        self.nucleus_sampling, self.nucleus_threshold = bool(nucleus_sampling), float(nucleus_threshold)
        if self.tilt_dict["tilt"]:
            self.goal_tilt, self.veh_veh_tilt = self.tilt_dict["goal_tilt"], self.tilt_dict["veh_veh_tilt"]
            self.veh_edge_tilt = self.tilt_dict["veh_edge_tilt"]
        self.seed = seed
        self.chunk_groups = chunk_groups
        self.lib = model.lib
        self.groups_last_step = 0
        self.launch_count = 0
        self._map_cache = None  # device memory of the per-focal polyline-encoder cache (steps 0..31)
        self.use_map_cache = os.environ.get("CTRLSIM_MAP_CACHE", "1") != "0"
        self._prefix_cache = None  # device memory of the decoder prefix cache (steps 0..31), one slot per chunk
        self._prefix_geom = None
        self.use_prefix_cache = os.environ.get("CTRLSIM_PREFIX_CACHE", "1") != "0"
        self.prefix_cache_max_bytes = int(float(os.environ.get("CTRLSIM_PREFIX_CACHE_GB", "96")) * 2**30)

    def _params(self):
        td = self.tilt_dict
        tilt = (C.c_double * 3)(float(td["goal_tilt"] or 0), float(td["veh_veh_tilt"] or 0), float(td["veh_edge_tilt"] or 0))
        return _lib.CtrlSimPolicyParams(seed=self.seed, tilt=tilt, temperature=float(self.action_temperature),
                                        tilt_enabled=1 if td["tilt"] else 0,
                                        nucleus_sampling=1 if self.nucleus_sampling else 0,
                                        nucleus_threshold=self.nucleus_threshold,
                                        rtg_mode=1 if self.real_time_rewards else 0)

    def _reward_params(self):
        # policy_evaluator.py:127-143: max_return wins over min_return
        mode = "max_return" if self.max_return else ("min_return" if self.min_return else "data")
        if not self.use_rtg:
            mode = "unused"
        return _lib.make_reward_params(self.cfg, mode)

    def _stream(self):
        return torch.cuda.current_stream(self.model.device).cuda_stream

    # ---- the reference's four verbs, batched -----------------------------------------------------------------------
    def attach_caches(self, batch: SceneBatch):
        """(Re)attach this policy's polyline-encoder cache, sized for ``batch``, to its model handle."""
        if self.use_map_cache:
            need = int(self.lib.ctrlsim_map_cache_bytes(batch.S, batch.N))
            if self._map_cache is None or self._map_cache.numel() < need:
                self._map_cache = torch.empty(need, dtype=torch.uint8, device=self.model.device)
            _lib.check(self.lib.ctrlsim_attach_map_cache(self.model.handle, self._map_cache.data_ptr(),
                                                         self._map_cache.numel()), "ctrlsim_attach_map_cache")
        else:
            _lib.check(self.lib.ctrlsim_attach_map_cache(self.model.handle, None, 0), "ctrlsim_attach_map_cache")

    def reset(self, batch: SceneBatch):
        if batch.S == 0:  # a rank that owns no scene (fewer scenes than ranks) only takes part in the collective
            return
        batch.reset_dynamic()
        self.attach_caches(batch)
        _lib.check(self.lib.ctrlsim_sim_reset(self.model.handle, batch.ptr, self._stream()), "ctrlsim_sim_reset")

    def update_state(self, batch: SceneBatch, t: int):
        """update_vehicle_data_dict + Policy.update_state for step t (observation, reward, history append)."""
        if batch.S == 0:
            return
        _lib.check(self.lib.ctrlsim_observe(self.model.handle, batch.ptr, t, self._stream()), "ctrlsim_observe")
        if self.real_time_rewards:  # compute_dense_reward + RTG bookkeeping (policy_evaluator.py:123-156)
            rp = self._reward_params()
            _lib.check(self.lib.ctrlsim_dense_reward(self.model.handle, batch.ptr, C.byref(rp), t, self._stream()),
                       "ctrlsim_dense_reward")

    def predict(self, batch: SceneBatch, t: int):
        """Focal grouping, tokenisation, two-pass network, RTG + action sampling; leaves next_action on the device."""
        if batch.S == 0:
            return 0
        st = self._stream()
        _lib.check(self.lib.ctrlsim_plan_groups(self.model.handle, batch.ptr, t, batch.n_total.data_ptr(), st),
                   "ctrlsim_plan_groups")
        n_total = int(batch.n_total.item())
        self.groups_last_step = n_total
        chunk = max(self.chunk_groups, batch.N)  # a scene's groups are processed together (RTG resolution)
        ws = self.model.workspace(chunk)
        if t == 0:
            self._attach_prefix_cache(chunk, n_total)
        p = self._params()
        _lib.check(self.lib.ctrlsim_policy_step(self.model.handle, batch.ptr, C.byref(p), t, n_total, ws.data_ptr(),
                                                ws.numel(), chunk, st), "ctrlsim_policy_step")
        return n_total

    def _attach_prefix_cache(self, chunk: int, n_total: int):
        """Size the prefix cache for this episode: one slot per chunk (scene runs pack a little worse than
        n_total / chunk), bounded by prefix_cache_max_bytes and by the free device memory."""
        if not self.use_prefix_cache or n_total <= 0:
            if self._prefix_geom is not None:
                _lib.check(self.lib.ctrlsim_attach_prefix_cache(self.model.handle, None, 0, 0), "ctrlsim_attach_prefix_cache")
                self._prefix_geom = None
            return
        per = int(self.lib.ctrlsim_prefix_cache_bytes(chunk, 1))
        want = -(-n_total // chunk) + 1 + n_total // (8 * chunk)
        have = 0 if self._prefix_cache is None else self._prefix_cache.numel()
        free, _ = torch.cuda.mem_get_info(self.model.device)
        budget = min(self.prefix_cache_max_bytes, have + max(0, free - (6 << 30)))
        slots = max(0, min(want, budget // per))
        if self._prefix_geom == (chunk, slots) and self._prefix_cache is not None:
            return  # the handle keeps its directory; entries are invalidated at t = 0 anyway
        if slots == 0:
            _lib.check(self.lib.ctrlsim_attach_prefix_cache(self.model.handle, None, 0, 0), "ctrlsim_attach_prefix_cache")
            self._prefix_geom = None
            return
        if have < slots * per:
            self._prefix_cache = None
            self._prefix_cache = torch.empty(slots * per, dtype=torch.uint8, device=self.model.device)
        _lib.check(self.lib.ctrlsim_attach_prefix_cache(self.model.handle, self._prefix_cache.data_ptr(), slots * per,
                                                        chunk), "ctrlsim_attach_prefix_cache")
        self._prefix_geom = (chunk, slots)

    def act(self, batch: SceneBatch, t: int):
        """policy.act / apply_gt_action for every vehicle, then Simulation.step(dt)."""
        if batch.S == 0:
            return
        _lib.check(self.lib.ctrlsim_sim_step(self.model.handle, batch.ptr, t, self._stream()), "ctrlsim_sim_step")


def _jsd(p, q):
    """scipy.spatial.distance.jensenshannon with the natural log (evaluators/policy_evaluator.py:269)."""
    p = np.asarray(p, np.float64)
    q = np.asarray(q, np.float64)
    p, q = p / p.sum(), q / q.sum()
    m = (p + q) / 2.0
    left = np.where(p > 0, p * np.log(np.where(p > 0, p, 1.0) / np.where(m > 0, m, 1.0)), 0.0)
    right = np.where(q > 0, q * np.log(np.where(q > 0, q, 1.0) / np.where(m > 0, m, 1.0)), 0.0)
    return float(np.sqrt((left.sum() + right.sum()) / 2.0))


EVAL_MODES = ("multi_agent", "one_agent", "two_agent")


class B200PolicyEvaluator:
    def __init__(self, cfg, policy: B200Policy, scenes=None, scene_ids=None):
        """``scenes`` (optional): in-memory list of {'json', 'preproc'} dicts; otherwise the reference's files are read:
        <dataset_root>/test_filenames.pkl, <nocturne_waymo_val_folder>/<name>.json and
        <dataset_root>/preprocess/test/<name>_physics.pkl (evaluators/policy_evaluator.py:33-41, evaluator.py:44-57)."""
        self.cfg, self.policy = cfg, policy
        self.cfg_rl_waymo = cfg.dataset.waymo
        self.steps, self.dt, self.history_steps = cfg.nocturne.steps, cfg.nocturne.dt, cfg.nocturne.history_steps
        self.rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        self.world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        # reading from disk: the reference stops after num_files_to_evaluate // partitions scenes that were actually
        # EVALUATED - scenes without preprocessed data or without a candidate agent do not count
        # (policy_evaluator.py:436-437,445-446,461-464,492), so files are read lazily while scenes are selected
        self._from_files = scenes is None
        self.scenes = [] if scenes is None else scenes
        self.scene_ids = list(range(len(self.scenes))) if scene_ids is None else list(scene_ids)
        self.scenes_presharded = False  # True: ``scenes`` already is this rank's share (no k mod world selection)
        self.batch = None
        self.last_summary = None
        self.contact_overflow = 0

    def _iter_files(self):
        with open(os.path.join(self.cfg.dataset_root, "test_filenames.pkl"), "rb") as f:
            names = pickle.load(f)["test_filenames"]
        for i, name in enumerate(names):
            pkl = os.path.join(self.cfg.dataset_root, "preprocess/test", f"{name[:-5]}_physics.pkl")
            if not os.path.exists(pkl):
                continue  # the reference silently skips scenes without preprocessed data (policy_evaluator.py:445-446)
            with open(os.path.join(self.cfg.nocturne_waymo_val_folder, name)) as f:
                js = json.load(f)
            with open(pkl, "rb") as f:
                pre = pickle.load(f)
            yield i, {"name": name, "json": js, "preproc": pre}

    def iter_selected(self, eval_threshold=None, keep_replay_only=False):
        """Walk the scenes in file order with the evaluation's seeded generator, draw the evaluated vehicles of each
        (policy_evaluator.py:450-464) and yield this rank's share (scene k -> rank k mod world, k counting accepted
        scenes) one at a time as (scene, id, parsed, evaluated_set) - so that a caller can overlap the host-side
        parsing of later scenes with the device work on earlier ones (evaluate_policy(sub_batch_scenes=...))."""
        from .scenario import interesting_pairs, parse_scenario
        cfg = self.cfg
        mode = cfg.eval.eval_mode
        if mode not in EVAL_MODES:
            raise ValueError(f"cfg.eval.eval_mode={mode!r}: expected one of {EVAL_MODES} (cfgs/eval/base.yaml:13-14)")
        rng = random.Random(cfg.eval.seed)
        thr = cfg.eval.multi_agent_eval_threshold if eval_threshold is None else eval_threshold
        sc = cfg.nocturne["scenario"]
        if self._from_files:
            source = self._iter_files()
            limit = cfg.eval.num_files_to_evaluate // cfg.eval.partitions
            self.scenes, self.scene_ids = [], []
        else:
            source = zip(self.scene_ids, self.scenes)
            limit = None
        accepted = 0
        for sid, s in source:
            if limit is not None and accepted == limit:
                break
            p = parse_scenario(s["json"], self.steps, sc["moving_threshold"], sc["speed_threshold"])
            moving = [i for i in range(p["n"]) if p["moving"][i]]
            if mode == "multi_agent":  # policy_evaluator.py:450-454
                ev = rng.sample(moving, thr) if len(moving) > thr else moving
            else:  # one_agent / two_agent: a random "interesting" pair, or its first vehicle (policy_evaluator.py:455-459)
                e = cfg.eval
                pairs = interesting_pairs(p, moving, self.history_steps, e.interesting_traj_len_threshold,
                                          e.interesting_goal_dist_threshold, e.interesting_timestep_diff_threshold)
                pick = rng.choice(pairs) if pairs else None  # the only use of the generator, only when a pair exists
                ev = [] if pick is None else ([pick[0]] if mode == "one_agent" else [pick[0], pick[1]])
            if not ev and not keep_replay_only:
                continue  # no candidate agent: scene skipped (policy_evaluator.py:461-464)
            own = self.scenes_presharded or accepted % self.world == self.rank
            accepted += 1
            if self._from_files:
                self.scenes.append(s if own else None)  # other ranks' scenes are not kept in memory
                self.scene_ids.append(sid)
            if own:
                yield s, sid, p, ev

    def select_scenes(self, eval_threshold=None, keep_replay_only=False):
        """Host half of build_batch: all of this rank's scenes at once. Returns (scenes, ids, parsed, evaluated_sets,
        threshold)."""
        thr = self.cfg.eval.multi_agent_eval_threshold if eval_threshold is None else eval_threshold
        mine, mine_ids, mine_parsed, mine_ev = [], [], [], []
        for s, sid, p, ev in self.iter_selected(eval_threshold, keep_replay_only):
            mine.append(s)
            mine_ids.append(sid)
            mine_parsed.append(p)
            mine_ev.append(ev)
        return mine, mine_ids, mine_parsed, mine_ev, thr

    def build_batch(self, eval_threshold=None, keep_replay_only=False):
        """Shard scenes over ranks keeping the evaluated-vehicle draw of the single-process evaluator: every rank walks
        all scenes in order with the same seeded generator (select_scenes).
        ``keep_replay_only``: keep scenes without any evaluated vehicle (pure log replay, BASELINE config 4) instead of
        skipping them like the reference evaluator does."""
        mine, mine_ids, mine_parsed, mine_ev, thr = self.select_scenes(eval_threshold, keep_replay_only)
        self.batch = SceneBatch(self.cfg, mine, mine_ids, self.policy.model.device, thr, parsed=mine_parsed,
                                evaluated_sets=mine_ev)
        return self.batch

    def rollout(self, batch=None, max_steps=None):
        """The 90-step closed loop of evaluate_policy (policy_evaluator.py:514-557) for the whole batch."""
        b = batch or self.batch
        pol = self.policy
        pol.reset(b)
        steps = self.steps if max_steps is None else max_steps
        for t in range(steps):
            pol.update_state(b, t)
            pol.predict(b, t)
            pol.act(b, t)
        if steps == self.steps:
            pol.update_state(b, self.steps)
        return b

    def summarize(self, batch=None, local_only=False):
        b = batch or self.batch
        dev = self.policy.model.device
        out_scene = torch.zeros(b.S, 8, dtype=torch.float64, device=dev)
        out_hist = torch.zeros(8, 200, dtype=torch.int64, device=dev)
        if b.S > 0:  # a rank without scenes contributes zeros to the all-reduce below
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(self.policy.lib.ctrlsim_metrics(self.policy.model.handle, b.ptr, out_scene.data_ptr(),
                                                       out_hist.data_ptr(), st), "ctrlsim_metrics")
            self.contact_overflow = b.contact_overflow()
            if self.contact_overflow:
                warnings.warn(f"{self.contact_overflow} contact(s) exceeded the simulator's capacity (128 broad-phase pairs "
                              "per scene / 32 contacts per island, sim_contacts.cuh) and were dropped: the affected "
                              "scenes no longer follow the reference's Box2D step")
        # flat summary: [goal_sum, n_agents, sum_scene_coll, sum_scene_off, n_scenes_with_agents, ade_sum, fde_sum, 0]
        summ = torch.cat([out_scene.sum(0), out_hist.to(torch.float64).flatten()])
        if local_only:  # a sub-batch of a pipelined evaluation: the caller adds the summaries up and reduces once
            return summ.cpu().numpy()
        if self.world > 1:
            torch.distributed.all_reduce(summ)  # the one collective of an evaluation
        self.last_summary = summ.cpu().numpy()
        return self.last_summary

    @staticmethod
    def metrics_from_summary(s, accel_bins=20):
        goal_sum, n_ag, coll_sum, off_sum, n_sc, ade_sum, fde_sum = s[:7]
        h = s[8:].reshape(8, 200)
        if n_ag == 0 or n_sc == 0:
            raise ValueError("no scene with an evaluated vehicle was accepted: there is nothing to report")
        m = {"goal": goal_sum / n_ag, "collision_rate": coll_sum / n_sc, "offroad_rate": off_sum / n_sc,
             "fde": fde_sum / n_ag, "ade": ade_sum / n_ag,
             "lin_speed_jsd": _jsd(h[0], h[1]), "ang_speed_jsd": _jsd(h[2], h[3]),
             "accel_jsd": _jsd(h[4][:accel_bins], h[5][:accel_bins]), "nearest_dist_jsd": _jsd(h[6], h[7])}
        return {k: float(v) for k, v in m.items()}

    def scene_results(self, batch=None):
        """The per-agent / per-sample lists the reference accumulates in update_running_statistics
        (policy_evaluator.py:162-248) and saves per partition (``saved_scene_metrics``, :578-583), for THIS rank's scenes,
        computed on the host from the device trace: goal_success, ade, fde (one entry per evaluated vehicle with at least
        one existing future step), collision / off_road (one per scene), and the raw samples behind the four JSD metrics."""
        b = batch or self.batch
        tr = b.trace()
        gt = b.t["gt"].cpu().numpy()
        T1, hist, dt = self.steps + 1, self.history_steps, self.dt
        out = {k: [] for k in ("goal_success", "ade", "fde", "accel_gt", "accel_sim", "ang_speed_gt", "ang_speed_sim",
                               "lin_speed_gt", "lin_speed_sim", "nearest_dist_gt", "nearest_dist_sim", "collision", "off_road")}
        for s in range(b.S):
            colls, offs = [], []
            for v in b.evaluated_ids[s]:
                mask = tr["tr_exist"][s, v].astype(bool).copy()
                mask[:hist] = False
                if mask.sum() == 0:
                    continue
                rew = tr["tr_reward"][s, v].astype(np.float64)[mask]
                out["goal_success"].append(float(np.any(rew[:, 0] == 1)))
                colls.append(float(np.any(rew[:, 6] == 1)))
                offs.append(float(np.any(rew[:, 7] == 1)))
                pos, gpos = tr["tr_pos"][s, v].astype(np.float64), gt[s, v, :, :2]
                out["ade"].append(float(np.linalg.norm(pos[mask] - gpos[mask], axis=1).mean()))
                last = np.where(mask)[0][-1]
                out["fde"].append(float(np.linalg.norm(pos[last] - gpos[last])))
                out["lin_speed_sim"].append(np.linalg.norm(tr["tr_vel"][s, v].astype(np.float64)[mask], axis=1))
                out["lin_speed_gt"].append(gt[s, v, :, 3][mask])
                out["ang_speed_sim"].append(tr["tr_heading"][s, v].astype(np.float64)[mask] / dt)
                out["ang_speed_gt"].append(gt[s, v, :, 2][mask] / dt)
                gacc = np.zeros(T1)  # central difference of the logged speed, 0 at both ends (policy_evaluator.py:107-111)
                gacc[1:self.steps - 1] = (gt[s, v, 2:self.steps, 3] - gt[s, v, 0:self.steps - 2, 3]) / (2 * dt)
                am = np.ones(int(mask.sum()), bool)
                am[0] = am[-1] = False
                out["accel_gt"].append(gacc[mask][am])
                out["accel_sim"].append(tr["tr_action"][s, v, :, 0][mask][am])
                out["nearest_dist_sim"].append(tr["tr_nearest"][s, v, :, 0][mask])
                out["nearest_dist_gt"].append(tr["tr_nearest"][s, v, :, 1][mask])
            if colls:
                out["collision"].append(float(np.mean(colls)))
                out["off_road"].append(float(np.mean(offs)))
        return out

    def write_partition_metrics(self, path=None, batch=None):
        """The per-partition JSON of the reference (policy_evaluator.py:578-593: ``<model dir>/scene_results/
        partition_<k>.json`` with the lists of scene_results(); the reference writes it for its partitioned CTG++ runs,
        ``eval=partitioned``) for this rank's scenes.  Returns the path."""
        res = self.scene_results(batch)
        ser = {k: [x.tolist() if isinstance(x, np.ndarray) else x for x in v] for k, v in res.items()}
        if path is None:
            part = self.cfg.eval.get("partition", self.rank) if hasattr(self.cfg.eval, "get") else self.rank
            out_dir = os.path.join(os.path.dirname(str(self.policy.model_path)), "scene_results")
            os.makedirs(out_dir, exist_ok=True)
            path = os.path.join(out_dir, f"partition_{part}.json")
        with open(path, "w") as f:
            json.dump(ser, f)
        return path

    def evaluate_policy(self, sub_batch_scenes=None, eval_threshold=None, keep_traces=False):
        """PolicyEvaluator.evaluate_policy() -> (metrics_dict, [str]).

        ``sub_batch_scenes``: evaluate this rank's scenes in sub-batches of that many scenes, with a host thread that
        parses the next sub-batch's scenario JSONs and uploads its arrays WHILE the device rolls the current one out
        (parsing is ~6 ms of Python per 64-vehicle scene; the device needs ~0.18 s per scene-episode, so everything but
        the first sub-batch's parse is hidden).  Scene results do not depend on how scenes are batched
        (test_full_size_batch_is_deterministic_and_shard_invariant) and the summary vectors add, so the metrics are
        those of the one-batch evaluation.  ``keep_traces``: keep every sub-batch's trace() in ``self.traces``."""
        if sub_batch_scenes is None:
            if self.batch is None:
                self.build_batch(eval_threshold)
            self.rollout()
            if keep_traces:
                self.traces = [self.batch.trace()]
            m = self.metrics_from_summary(self.summarize(), self.cfg_rl_waymo.accel_discretization)
            return m, ["{}: {:.6f}".format(k, v) for k, v in m.items()]
        import queue
        import threading
        thr = self.cfg.eval.multi_agent_eval_threshold if eval_threshold is None else eval_threshold
        dev = self.policy.model.device
        q: "queue.Queue" = queue.Queue(maxsize=2)

        on_gpu = torch.device(dev).type == "cuda"
        side = torch.cuda.Stream(device=dev) if on_gpu else None  # uploads must not queue behind the rollout's kernels

        def upload(cur):
            if side is None:
                return SceneBatch(self.cfg, cur[0], cur[1], dev, thr, parsed=cur[2], evaluated_sets=cur[3])
            with torch.cuda.stream(side):
                b = SceneBatch(self.cfg, cur[0], cur[1], dev, thr, parsed=cur[2], evaluated_sets=cur[3])
            side.synchronize()  # the consumer uses the arrays on its own stream
            return b

        stop = threading.Event()  # set by the consumer when it gives up: the producer must not block on a full queue

        def hand_over(x):
            while not stop.is_set():
                try:
                    q.put(x, timeout=0.2)
                    return True
                except queue.Full:
                    pass
            return False

        def producer():
            try:
                cur = ([], [], [], [])
                for item in self.iter_selected(eval_threshold):
                    if stop.is_set():
                        return
                    for lst, v in zip(cur, item):
                        lst.append(v)
                    if len(cur[0]) == sub_batch_scenes:
                        if not hand_over(upload(cur)):
                            return
                        cur = ([], [], [], [])
                if cur[0] and not hand_over(upload(cur)):
                    return
                hand_over(None)
            except BaseException as e:  # noqa: BLE001 - handed to the consumer, which re-raises it
                hand_over(e)

        th = threading.Thread(target=producer, daemon=True)
        th.start()
        total, self.traces, self.n_evaluated, overflow = None, [], 0, 0
        try:
            while True:
                b = q.get()
                if b is None:
                    break
                if isinstance(b, BaseException):
                    raise b
                self.batch = b
                self.rollout(b)
                sm = self.summarize(b, local_only=True)
                overflow += self.contact_overflow
                total = sm if total is None else total + sm
                self.n_evaluated += b.n_evaluated()
                if keep_traces:
                    self.traces.append(b.trace())
        finally:
            stop.set()  # an error on this side (or the normal end) releases a producer waiting on the queue
            th.join()
        self.contact_overflow = overflow
        summ = torch.as_tensor(total if total is not None else np.zeros(8 + 8 * 200), dtype=torch.float64, device=dev)
        if self.world > 1:
            torch.distributed.all_reduce(summ)  # the one collective of an evaluation
        self.last_summary = summ.cpu().numpy()
        m = self.metrics_from_summary(self.last_summary, self.cfg_rl_waymo.accel_discretization)
        return m, ["{}: {:.6f}".format(k, v) for k, v in m.items()]
