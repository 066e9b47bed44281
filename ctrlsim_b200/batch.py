"""GPU-resident scenario batch (struct of arrays in HBM) backing struct CtrlSimBatch of include/ctrlsim_b200.h."""
from __future__ import annotations

import ctypes as C
import random

import numpy as np
import torch

from . import lib as _lib
from .scenario import initial_rtgs, parse_scenario, road_arrays, road_edge_polylines

# per-policy state: everything a Policy object owns in the reference (policies/policy.py:45-59: buffers, relevant agent
# sets) plus the focal groups of the current step; a policy_view() gets its own copies, everything else is the shared world
POLICY_FIELDS = ("evaluated", "eval_order", "hist_rtg", "relevant", "next_action", "tr_rtg_idx", "tr_act_idx", "n_groups",
                 "group_off", "group_focal", "group_members", "group_served", "group_scene", "group_local", "rt_rtg")
STATIC_FIELDS = {"scene_id", "n_veh", "veh_len", "veh_wid", "gt", "gt_valid", "goal", "goal_norm", "evaluated",
                 "eval_order", "road_xy", "road_valid", "road_type", "n_poly", "segs", "n_seg",
                 "edge_xy", "edge_off", "n_edge", "rtg_init"}

_DT = {"int64": torch.int64, "int32": torch.int32, "int16": torch.int16, "int8": torch.int8, "uint8": torch.uint8,
       "float32": torch.float32, "float64": torch.float64}
MAX_VEH = 64


class SceneBatch:
    """S scenes padded to common (N vehicles, Pm polylines, E edge segments).

    ``scenes``: list of dicts with ``json`` (Nocturne schema) and ``preproc`` (road_points / road_types), e.g. from
    ctrlsim_b200.synth.make_scene or read from the reference's files.  ``scene_ids`` are the global indices used by the
    sampler (so sharding scenes over ranks does not change any draw).  ``rng`` reproduces the evaluator's
    ``random.sample`` of evaluated vehicles (evaluators/policy_evaluator.py:450-454); pass the same ``random.Random``
    across shards in scene order.
    """

    def __init__(self, cfg, scenes, scene_ids=None, device="cuda:0", eval_threshold=None, rng=None, parsed=None,
                 evaluated_sets=None):
        self.cfg = cfg
        self.device = torch.device(device)
        steps = cfg.nocturne.steps
        self.steps = steps
        thr = cfg.eval.multi_agent_eval_threshold if eval_threshold is None else eval_threshold
        rng = rng or random.Random(cfg.eval.seed)
        sc_cfg = cfg.nocturne["scenario"]
        if parsed is None:
            parsed = [parse_scenario(s["json"], steps, sc_cfg["moving_threshold"], sc_cfg["speed_threshold"]) for s in scenes]
        roads = [road_arrays(s["preproc"]) for s in scenes]
        # real-time rewards: the road-edge polylines the signed distance is measured to, and the RTGs of
        # the logged episode at t = 0 when the *_physics.pkl carries what they are derived from
        # (a scene handed over pre-parsed, without its JSON, may carry them as 'edge_polylines'; else it has none)
        edges = [p["edge_polylines"] if "edge_polylines" in p else
                 (road_edge_polylines(s["json"]) if "json" in s else list(s.get("edge_polylines", [])))
                 for s, p in zip(scenes, parsed)]
        S = len(scenes)
        N = max(1, max((p["n"] for p in parsed), default=0))  # S == 0: a rank without scenes (more ranks than scenes)
        if N > MAX_VEH:
            raise ValueError(f"at most {MAX_VEH} vehicles per scene are supported (got {N})")
        Pm = max(1, max((r[0].shape[0] for r in roads), default=0))
        E = max(1, max((p["segs"].shape[0] for p in parsed), default=0))
        T1 = steps + 1
        Ep = max(1, max((sum(len(q) for q in e) for e in edges), default=0))
        Pe = max(1, max((len(e) for e in edges), default=0))
        dims = {"S": S, "N": N, "Pm": Pm, "E": E, "T": steps, "T1": T1, "A": cfg.dataset.waymo.max_num_agents,
                "Ep": Ep, "Pe": Pe}
        self.S, self.N, self.Pm, self.E, self.Ep, self.Pe = S, N, Pm, E, Ep, Pe
        host = {}
        for name, dt, shp in _lib.BATCH_FIELDS:
            shape = [eval(tok, {}, dims) for tok in shp.split(",")]
            host[name] = np.zeros(shape, dtype=dt)
        host["eval_order"][:] = -1
        host["road_type"][:] = -1
        host["hist_rtg"][..., 1:] = 35  # un-sampled RTG (0,0,0) re-normalises to bins (0,35,35)
        host["tr_rtg_idx"][:] = -1
        host["tr_act_idx"][:] = -1
        self.evaluated_ids = []
        self._gt_len = [p["gt_valid"].astype(np.float64).sum(axis=1).astype(np.int64) for p in parsed]
        ids = list(range(S)) if scene_ids is None else list(scene_ids)
        for s, (p, (rxy, rvalid, rtype)) in enumerate(zip(parsed, roads)):
            n = p["n"]
            host["scene_id"][s] = ids[s]
            host["n_veh"][s] = n
            host["veh_len"][s, :n], host["veh_wid"][s, :n] = p["size"][:, 0], p["size"][:, 1]
            host["gt"][s, :n], host["gt_valid"][s, :n] = p["gt"], p["gt_valid"]
            host["goal"][s, :n], host["goal_norm"][s, :n] = p["goal"], p["goal_norm"]
            moving = [i for i in range(n) if p["moving"][i]]
            if evaluated_sets is not None:
                ev = list(evaluated_sets[s])
            else:
                ev = rng.sample(moving, thr) if len(moving) > thr else moving
            self.evaluated_ids.append(sorted(ev))
            host["evaluated"][s, ev] = 1
            if ev:  # descending GT length, exactly the reference's np.argsort(...)[::-1] (autoregressive_policy.py:88-94)
                lengths = [int(p["gt_valid"][v].astype(np.float64).sum()) for v in ev]
                order = np.argsort(np.array(lengths))[::-1]
                host["eval_order"][s, :len(ev)] = np.array(ev)[order]
            np_ = rxy.shape[0]
            host["n_poly"][s] = np_
            host["road_xy"][s, :np_], host["road_valid"][s, :np_], host["road_type"][s, :np_] = rxy, rvalid, rtype
            ns = p["segs"].shape[0]
            host["n_seg"][s] = ns
            host["segs"][s, :ns] = p["segs"]
            host["n_edge"][s] = len(edges[s])
            o = 0
            for k, q in enumerate(edges[s]):
                host["edge_off"][s, k] = o
                host["edge_xy"][s, o:o + len(q)] = q
                o += len(q)
            host["edge_off"][s, len(edges[s]):] = o
            pre = scenes[s]["preproc"]
            if all(k in pre for k in ("ag_data", "ag_rewards", "veh_edge_dist_rewards", "veh_veh_dist_rewards")) \
                    and len(pre["ag_data"]) == n:
                host["rtg_init"][s, :n] = initial_rtgs(cfg, pre)
        self.t = {k: torch.from_numpy(v).to(self.device) for k, v in host.items()}
        self._init_dynamic = {k: host[k] for k in ("hist_rtg", "tr_rtg_idx", "tr_act_idx")}
        self.struct = _lib.CtrlSimBatch(n_scenes=S, max_veh=N, max_poly=Pm, max_seg=E, max_edge_pts=Ep, max_edge_poly=Pe,
                                        **{k: self.t[k].data_ptr() for k, _, _ in _lib.BATCH_FIELDS})
        self.n_total = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._owned = set(self.t)

    def policy_view(self, evaluated_sets, gt=None):
        """A second policy's view of the SAME world (planner-vs-adversary evaluation, evaluators/
        planner_adversary_evaluator.py:466-540: two Policy objects share one vehicle_data_dict and one simulator).
        The view shares every world array (static scene data, simulator state, history of states and actions, trace)
        with this batch and owns fresh per-policy state (POLICY_FIELDS) for its own set of controlled vehicles.
        ``gt``: optional replacement of the log-replay targets [S,N,T1,4] (a scripted trajectory for some vehicle)."""
        v = object.__new__(SceneBatch)
        v.__dict__.update({k: getattr(self, k) for k in ("cfg", "device", "steps", "S", "N", "Pm", "E", "Ep", "Pe", "_gt_len")})
        v.t = dict(self.t)
        own = {}
        for k in POLICY_FIELDS:
            own[k] = torch.zeros_like(self.t[k])
        own["eval_order"].fill_(-1)
        own["hist_rtg"][..., 1:] = 35
        own["tr_rtg_idx"].fill_(-1)
        own["tr_act_idx"].fill_(-1)
        ev_host = np.zeros((self.S, self.N), np.uint8)
        order_host = -np.ones((self.S, self.N), np.int32)
        v.evaluated_ids = []
        for s, ev in enumerate(evaluated_sets):
            ev = [int(x) for x in ev]
            v.evaluated_ids.append(sorted(ev))
            ev_host[s, ev] = 1
            if ev:  # descending GT length like SceneBatch.__init__ (autoregressive_policy.py:88-94)
                order = np.argsort(np.array([int(self._gt_len[s][x]) for x in ev]))[::-1]
                order_host[s, :len(ev)] = np.array(ev)[order]
        own["evaluated"].copy_(torch.from_numpy(ev_host))
        own["eval_order"].copy_(torch.from_numpy(order_host))
        v._init_dynamic = {k: own[k].cpu().numpy() for k in ("hist_rtg", "tr_rtg_idx", "tr_act_idx")}
        v.t.update(own)
        v._owned = set(POLICY_FIELDS)
        if gt is not None:
            v.t["gt"] = torch.as_tensor(np.ascontiguousarray(gt, np.float64)).to(self.device)
            assert v.t["gt"].shape == self.t["gt"].shape
            v._owned.add("gt")
        v.struct = _lib.CtrlSimBatch(n_scenes=self.S, max_veh=self.N, max_poly=self.Pm, max_seg=self.E,
                                     max_edge_pts=self.Ep, max_edge_poly=self.Pe,
                                     **{k: v.t[k].data_ptr() for k, _, _ in _lib.BATCH_FIELDS})
        v.n_total = torch.zeros(1, dtype=torch.int32, device=self.device)
        return v

    @property
    def ptr(self):
        return C.byref(self.struct)

    def reset_dynamic(self):
        """Policy.reset (policies/policy.py:45-59): clear history, trace and group state for a new episode."""
        for k, v in self.t.items():
            if k in STATIC_FIELDS or k not in self._owned:
                continue
            if k in self._init_dynamic:
                v.copy_(torch.from_numpy(self._init_dynamic[k]))
            else:
                v.zero_()

    def contact_overflow(self) -> int:
        """Broad-phase pairs / island contacts the simulator had to drop since the last reset because a capacity of
        sim_contacts.cuh (128 pairs per scene, 32 contacts per island) was exceeded; > 0 means the episode is no longer
        the reference's (Box2D has no such caps)."""
        if self.S == 0:
            return 0
        return int(self.t["cstate"][:, 3].contiguous().view(torch.int32).sum().item())

    def n_evaluated(self) -> int:
        return sum(len(e) for e in self.evaluated_ids)

    def trace(self) -> dict:
        """Device -> host copy of everything the reference keeps in vehicle_data_dict (for parity dumps)."""
        keys = ("tr_pos", "tr_vel", "tr_heading", "tr_exist", "tr_action", "tr_reward", "tr_nearest", "tr_rtg_idx",
                "tr_act_idx", "n_veh", "tr_dense", "rt_rtg")
        return {k: self.t[k].cpu().numpy() for k in keys}
