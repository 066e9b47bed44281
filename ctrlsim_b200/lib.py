"""ctypes binding of include/ctrlsim_b200.h (the C-ABI of the CUDA extension).

The product path has no CPU fallback: if the shared library is missing it is built (nvcc must be present); if it
cannot be loaded, importing a GPU entry point raises.  PyTorch only provides device memory and streams; all pointers
handed over are ``tensor.data_ptr()`` values.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libctrlsim_b200.so")
LIB_PATH_WIDE = os.path.join(HERE, "lib", "libctrlsim_b200_wide.so")  # -DCTRLSIM_WIDE: 64 agents / 256 polylines per group
ABI_VERSION = 6
GEOMETRY = {False: (24, 200), True: (64, 256)}  # (max_num_agents, max_num_road_polylines) of the two builds


class CtrlSimConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("hidden_dim", C.c_int32), ("num_heads", C.c_int32), ("dim_feedforward", C.c_int32),
        ("enc_layers", C.c_int32), ("dec_layers", C.c_int32),
        ("max_agents", C.c_int32), ("context_len", C.c_int32), ("max_polylines", C.c_int32),
        ("pts_per_polyline", C.c_int32),
        ("n_action_bins", C.c_int32), ("n_steer_bins", C.c_int32), ("n_rtg_bins", C.c_int32),
        ("steps", C.c_int32), ("history_steps", C.c_int32),
        ("dt", C.c_float),
        ("agent_dist_threshold", C.c_double),
        ("min_accel", C.c_double), ("max_accel", C.c_double), ("min_steer", C.c_double), ("max_steer", C.c_double),
        ("pos_tol", C.c_double), ("heading_tol", C.c_double), ("speed_tol", C.c_double),
        ("goal_dist_scaling", C.c_double), ("reward_scaling", C.c_double),
        ("decision_transformer", C.c_int32), ("reserved0", C.c_int32),
        ("rtg_min", C.c_double * 3), ("rtg_max", C.c_double * 3),
    ]


# (field name, torch dtype name, shape expression) in the exact order of struct CtrlSimBatch; A = max_agents of the
# library the batch is stepped with (24, or 64 for the wide build)
BATCH_FIELDS = [
    ("scene_id", "int64", "S"), ("n_veh", "int32", "S"), ("veh_len", "float32", "S,N"), ("veh_wid", "float32", "S,N"),
    ("gt", "float64", "S,N,T1,4"), ("gt_valid", "uint8", "S,N,T1"), ("goal", "float64", "S,N,4"),
    ("goal_norm", "float64", "S,N"), ("evaluated", "uint8", "S,N"), ("eval_order", "int32", "S,N"),
    ("road_xy", "float64", "S,Pm,100,2"), ("road_valid", "uint8", "S,Pm,100"), ("road_type", "int8", "S,Pm"),
    ("n_poly", "int32", "S"), ("segs", "float32", "S,E,4"), ("n_seg", "int32", "S"),
    ("body", "float32", "S,16,N"), ("obj", "float32", "S,4,N"), ("coll", "uint8", "S,2,N"),
    ("hist_state", "float64", "S,N,T,8"), ("hist_action", "float64", "S,N,T,2"), ("hist_rtg", "int16", "S,N,T,3"),
    ("relevant", "int64", "S,N"), ("next_action", "float64", "S,N,2"),
    ("tr_pos", "float32", "S,N,T1,2"), ("tr_vel", "float32", "S,N,T1,2"), ("tr_heading", "float32", "S,N,T1"),
    ("tr_exist", "uint8", "S,N,T1"), ("tr_action", "float64", "S,N,T1,2"), ("tr_reward", "float32", "S,N,T1,8"),
    ("tr_nearest", "float64", "S,N,T1,2"), ("tr_rtg_idx", "int16", "S,N,T,3"), ("tr_act_idx", "int16", "S,N,T"),
    ("n_groups", "int32", "S"), ("group_off", "int32", "S+1"), ("group_focal", "int32", "S,N"),
    ("group_members", "int32", "S,N,A"), ("group_served", "int64", "S,N"), ("group_scene", "int32", "S*N"),
    ("group_local", "int32", "S*N"), ("cstate", "float32", "S,4+8*N+20*128"),
    # ABI 6: real-time rewards (road-edge polylines for the signed distance, tracked RTG series, dense reward trace)
    ("edge_xy", "float64", "S,Ep,2"), ("edge_off", "int32", "S,Pe+1"), ("n_edge", "int32", "S"),
    ("rtg_init", "float64", "S,N,3"), ("rt_rtg", "float64", "S,N,T,3"), ("tr_dense", "float64", "S,N,T1,3"),
]


class CtrlSimBatch(C.Structure):
    _fields_ = ([("n_scenes", C.c_int32), ("max_veh", C.c_int32), ("max_poly", C.c_int32), ("max_seg", C.c_int32)]
                + [(name, C.c_void_p) for name, _, _ in BATCH_FIELDS]
                + [("max_edge_pts", C.c_int32), ("max_edge_poly", C.c_int32)])


class CtrlSimPolicyParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("tilt", C.c_double * 3), ("temperature", C.c_float),
                ("tilt_enabled", C.c_int32), ("nucleus_sampling", C.c_int32), ("nucleus_threshold", C.c_double),
                ("rtg_mode", C.c_int32), ("reserved0", C.c_int32)]


class CtrlSimRewardParams(C.Structure):
    _fields_ = [("max_veh_veh_distance", C.c_double), ("dist_to_road_edge_scaling_factor", C.c_double),
                ("veh_veh_collision_rew_multiplier", C.c_double), ("veh_edge_collision_rew_multiplier", C.c_double),
                ("pos_goal_shaped_min", C.c_double), ("pos_goal_shaped_max", C.c_double),
                ("pos_target_achieved_rew_multiplier", C.c_double),
                ("remove_shaped_goal", C.c_int32), ("remove_shaped_veh_reward", C.c_int32),
                ("remove_shaped_edge_reward", C.c_int32), ("return_mode", C.c_int32)]


class CtrlSimError(RuntimeError):
    pass


_lib = {}

_SIGS = {
    "ctrlsim_last_error": (C.c_char_p, []),
    "ctrlsim_abi_version": (C.c_int, []),
    "ctrlsim_create": (C.c_int, [C.POINTER(CtrlSimConfig), C.POINTER(C.c_void_p)]),
    "ctrlsim_destroy": (None, [C.c_void_p]),
    "ctrlsim_load_weights": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "ctrlsim_finalize_weights": (C.c_int, [C.c_void_p]),
    "ctrlsim_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int32]),
    "ctrlsim_map_cache_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "ctrlsim_attach_map_cache": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "ctrlsim_map_cache_stats": (None, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ctrlsim_prefix_cache_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "ctrlsim_attach_prefix_cache": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]),
    "ctrlsim_prefix_cache_stats": (None, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ctrlsim_sim_reset": (C.c_int, [C.c_void_p, C.POINTER(CtrlSimBatch), C.c_void_p]),
    "ctrlsim_observe": (C.c_int, [C.c_void_p, C.POINTER(CtrlSimBatch), C.c_int32, C.c_void_p]),
    "ctrlsim_dense_reward": (C.c_int, [C.c_void_p, C.POINTER(CtrlSimBatch), C.POINTER(CtrlSimRewardParams), C.c_int32,
                                       C.c_void_p]),
    "ctrlsim_plan_groups": (C.c_int, [C.c_void_p, C.POINTER(CtrlSimBatch), C.c_int32, C.c_void_p, C.c_void_p]),
    "ctrlsim_policy_step": (C.c_int, [C.c_void_p, C.POINTER(CtrlSimBatch), C.POINTER(CtrlSimPolicyParams), C.c_int32,
                                      C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "ctrlsim_sim_step": (C.c_int, [C.c_void_p, C.POINTER(CtrlSimBatch), C.c_int32, C.c_void_p]),
    "ctrlsim_metrics": (C.c_int, [C.c_void_p, C.POINTER(CtrlSimBatch), C.c_void_p, C.c_void_p, C.c_void_p]),
    "ctrlsim_launch_count": (C.c_longlong, []),
    "ctrlsim_debug_gemm": (None, [C.c_int32]),
    "ctrlsim_debug_gemm_trace": (None, [C.c_void_p]),
    "ctrlsim_debug_attn": (None, [C.c_int32]),
    "ctrlsim_debug_attn_trace": (None, [C.c_void_p]),
    "ctrlsim_profile_enable": (None, [C.c_int32]),
    "ctrlsim_profile_read": (None, [C.POINTER(C.c_double)]),
    "ctrlsim_linear": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 4 + [C.c_void_p]),
    "ctrlsim_layernorm": (C.c_int, [C.c_void_p] * 5 + [C.c_int32] * 2 + [C.c_void_p]),
    "ctrlsim_linear_res_ln": (C.c_int, [C.c_void_p] * 7 + [C.c_int32] * 2 + [C.c_void_p]),
    "ctrlsim_attn_padded": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "ctrlsim_attn_causal": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "ctrlsim_attn_causal_order": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "ctrlsim_attn_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_void_p]),
    "ctrlsim_attn_step_order": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "ctrlsim_map_pool": (C.c_int, [C.c_void_p] * 5 + [C.c_int32, C.c_void_p]),
    "ctrlsim_map_encode_pool": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "ctrlsim_sample_rows": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "ctrlsim_sample_rows_nucleus": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_void_p,
                                              C.c_double, C.c_void_p, C.c_void_p]),
    "ctrlsim_forward_tokens": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 11
                               + [C.c_void_p, C.c_int64, C.c_void_p]),
    "ctrlsim_forward_tokens_dt": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 9
                                  + [C.c_void_p, C.c_int64, C.c_void_p]),
    "ctrlsim_geom_poly_poly": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ctrlsim_geom_poly_seg": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
}
EXPORTS = tuple(_SIGS)


def is_wide(cfg) -> bool:
    """Which build serves ``cfg``: the reference-default geometry or the wide one; anything else is an error."""
    w = cfg.dataset.waymo
    geom = (w.max_num_agents, w.max_num_road_polylines)
    for wide, g in GEOMETRY.items():
        if geom == g:
            return wide
    raise CtrlSimError(f"no build of the library serves max_num_agents={geom[0]} / max_num_road_polylines={geom[1]}; "
                       f"available: {sorted(GEOMETRY.values())}")


def load(build_if_missing: bool = True, wide: bool = False):
    """Load libctrlsim_b200[_wide].so and declare every prototype of include/ctrlsim_b200.h."""
    if wide in _lib:
        return _lib[wide]
    path = LIB_PATH_WIDE if wide else LIB_PATH
    if not os.path.exists(path):
        if not build_if_missing:
            raise CtrlSimError(f"{path} is missing; run `python -m ctrlsim_b200.build`")
        from .build import build
        build(wide=wide)
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.ctrlsim_abi_version() != ABI_VERSION:
        raise CtrlSimError(f"{os.path.basename(path)} ABI version mismatch; rebuild with `python -m ctrlsim_b200.build -f`")
    _lib[wide] = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        # the message is thread-local inside the library that failed; ask every loaded build
        msg = b"; ".join(m for m in (l.ctrlsim_last_error() for l in _lib.values()) if m) if _lib else load().ctrlsim_last_error()
        raise CtrlSimError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


# switches of cfgs/config.yaml `nocturne:` whose reference default is compiled into observe_kernel / sim_step_kernel
# (utils/sim.py compute_reward, policy_evaluator.py collision handling); any other value must fail loudly, not silently
# evaluate something else
_FIXED_SWITCHES = {"collision_fix": True}
# (utils/sim.py:83-141 reads only these; shared_reward / collision_penalty / goal_distance_penalty belong to the RL env)
_FIXED_REWARD_SWITCHES = {"shaped_goal_distance": True, "position_target": True, "speed_target": True,
                          "heading_target": True}


def check_fixed_switches(cfg):
    n = cfg.nocturne
    bad = [f"nocturne.{k}={n[k]!r} (supported: {v!r})" for k, v in _FIXED_SWITCHES.items() if k in n and n[k] != v]
    rc = n["rew_cfg"]
    bad += [f"nocturne.rew_cfg.{k}={rc[k]!r} (supported: {v!r})" for k, v in _FIXED_REWARD_SWITCHES.items()
            if k in rc and rc[k] != v]
    if bad:
        raise NotImplementedError("the simulator / reward kernels implement the reference defaults only: " + "; ".join(bad))


def make_config(cfg) -> CtrlSimConfig:
    w, m, n = cfg.dataset.waymo, cfg.model, cfg.nocturne
    rc = n["rew_cfg"]
    check_fixed_switches(cfg)
    return CtrlSimConfig(
        abi_version=ABI_VERSION, hidden_dim=m.hidden_dim, num_heads=m.num_heads, dim_feedforward=m.dim_feedforward,
        enc_layers=m.num_transformer_encoder_layers, dec_layers=m.num_decoder_layers,
        max_agents=w.max_num_agents, context_len=w.train_context_length, max_polylines=w.max_num_road_polylines,
        pts_per_polyline=w.max_num_road_pts_per_polyline,
        n_action_bins=w.accel_discretization * w.steer_discretization, n_steer_bins=w.steer_discretization,
        n_rtg_bins=w.rtg_discretization, steps=n.steps, history_steps=n.history_steps, dt=n.dt,
        agent_dist_threshold=w.agent_dist_threshold, min_accel=w.min_accel, max_accel=w.max_accel,
        min_steer=w.min_steer, max_steer=w.max_steer, pos_tol=rc["position_target_tolerance"],
        heading_tol=rc["heading_target_tolerance"], speed_tol=rc["speed_target_tolerance"],
        goal_dist_scaling=rc.get("shaped_goal_distance_scaling", 1.0), reward_scaling=rc["reward_scaling"],
        decision_transformer=1 if m.get("decision_transformer", False) else 0,
        rtg_min=(C.c_double * 3)(w.min_rtg_pos, w.min_rtg_veh, w.min_rtg_road),
        rtg_max=(C.c_double * 3)(w.max_rtg_pos, w.max_rtg_veh, w.max_rtg_road))


RETURN_MODES = {"data": 0, "max_return": 1, "min_return": 2, "unused": 3}


def make_reward_params(cfg, return_mode: str = "data") -> CtrlSimRewardParams:
    """Constants of the real-time dense reward (cfgs/dataset/waymo/base.yaml:18-25,47-49) and how the RTG series starts
    (policy_evaluator.py:124-143: from the data, at the maximum return, or at the minimum one for evaluated vehicles)."""
    w = cfg.dataset.waymo
    return CtrlSimRewardParams(
        max_veh_veh_distance=w.max_veh_veh_distance, dist_to_road_edge_scaling_factor=w.dist_to_road_edge_scaling_factor,
        veh_veh_collision_rew_multiplier=w.veh_veh_collision_rew_multiplier,
        veh_edge_collision_rew_multiplier=w.veh_edge_collision_rew_multiplier,
        pos_goal_shaped_min=w.pos_goal_shaped_min, pos_goal_shaped_max=w.pos_goal_shaped_max,
        pos_target_achieved_rew_multiplier=w.pos_target_achieved_rew_multiplier,
        remove_shaped_goal=int(bool(w.remove_shaped_goal)), remove_shaped_veh_reward=int(bool(w.remove_shaped_veh_reward)),
        remove_shaped_edge_reward=int(bool(w.remove_shaped_edge_reward)), return_mode=RETURN_MODES[return_mode])
