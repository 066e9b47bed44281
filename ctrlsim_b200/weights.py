"""Deterministic synthetic weights in the reference's state-dict layout.

No trained checkpoint is available offline (the reference's weights live on Google Drive, README.md:47,108), so
benchmarks and parity tests use random-init weights of the exact architecture.  The layout (key names and shapes)
follows the reference modules: modules/encoder.py:18-46, modules/map_encoder.py:16-24, modules/decoder.py:16-27,
utils/layers.py:10-15.  Values come from a counter-based generator (philox.py) so that the same tensors can be
rebuilt anywhere without shipping a 33 MB file; magnitudes follow utils/train_utils.py:14-79 (xavier-uniform for
linear / attention projections, N(0, 0.02)-scale embeddings) except that biases and LayerNorm affines are made
non-trivial on purpose, so that a parity test catches a dropped bias or a swapped gamma/beta.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np

from .philox import uniform01


def _mlp(prefix, din, dh, dout):
    return [
        (f"{prefix}.mlp.0.weight", (dh, din), "linear"), (f"{prefix}.mlp.0.bias", (dh,), "bias"),
        (f"{prefix}.mlp.1.weight", (dh,), "ln_w"), (f"{prefix}.mlp.1.bias", (dh,), "ln_b"),
        (f"{prefix}.mlp.3.weight", (dout, dh), "linear"), (f"{prefix}.mlp.3.bias", (dout,), "bias"),
    ]


def _mha(prefix, h):
    return [
        (f"{prefix}.in_proj_weight", (3 * h, h), "inproj"), (f"{prefix}.in_proj_bias", (3 * h,), "bias"),
        (f"{prefix}.out_proj.weight", (h, h), "linear"), (f"{prefix}.out_proj.bias", (h,), "bias"),
    ]


def _ln(prefix, h):
    return [(f"{prefix}.weight", (h,), "ln_w"), (f"{prefix}.bias", (h,), "ln_b")]


def _lin(prefix, din, dout):
    return [(f"{prefix}.weight", (dout, din), "linear"), (f"{prefix}.bias", (dout,), "bias")]


def param_spec(cfg):
    """[(key, shape, kind)] in the reference's state_dict order."""
    m, w = cfg.model, cfg.dataset.waymo
    H, FF = m.hidden_dim, m.dim_feedforward
    n_act = w.accel_discretization * w.steer_discretization
    spec = []
    me = "encoder.map_encoder"
    spec += [(f"{me}.map_seeds", (1, 1, H), "seed")]
    spec += _mlp(f"{me}.road_pts_encoder", m.map_attr, H, H)
    spec += _mha(f"{me}.road_pts_attn_layer", H)
    spec += _ln(f"{me}.norm1", H) + _ln(f"{me}.norm2", H)
    spec += _mlp(f"{me}.map_feats", H, H, H)
    spec += _mlp(f"{me}.road_type_encoder", m.num_road_types, H, H)
    spec += _mlp(f"{me}.road_road_type_encoder", 2 * H, H, H)
    spec += _mlp("encoder.embed_state", m.state_dim, H, H)
    spec += _mlp("encoder.embed_goal", w.goal_dim, H, H)
    spec += _lin("encoder.embed_state_goal", 2 * H, H)
    spec += [("encoder.embed_action.weight", (n_act, H), "emb")]
    for c in ("goal", "veh", "road"):
        if m.get("decision_transformer", False):  # continuous RTG inputs: nn.Linear(1, H) (modules/encoder.py:27-30)
            spec += _lin(f"encoder.embed_rtg_{c}", 1, H)
        else:
            spec += [(f"encoder.embed_rtg_{c}.weight", (w.rtg_discretization, H), "emb")]
    spec += _lin("encoder.embed_rtg", H * m.num_reward_components, H)
    spec += [("encoder.embed_timestep.weight", (w.max_timestep, H), "emb"),
             ("encoder.embed_agent_id.weight", (w.max_num_agents, H), "emb")]
    spec += _ln("encoder.embed_ln", H)
    for l in range(m.num_transformer_encoder_layers):
        p = f"encoder.transformer_encoder.layers.{l}"
        spec += _mha(f"{p}.self_attn", H) + _lin(f"{p}.linear1", H, FF) + _lin(f"{p}.linear2", FF, H)
        spec += _ln(f"{p}.norm1", H) + _ln(f"{p}.norm2", H)
    for l in range(m.num_decoder_layers):
        p = f"decoder.transformer_decoder.layers.{l}"
        spec += _mha(f"{p}.self_attn", H) + _mha(f"{p}.multihead_attn", H)
        spec += _lin(f"{p}.linear1", H, FF) + _lin(f"{p}.linear2", FF, H)
        spec += _ln(f"{p}.norm1", H) + _ln(f"{p}.norm2", H) + _ln(f"{p}.norm3", H)
    spec += _mlp("decoder.predict_action", H, H, n_act)
    if m.predict_rtg:  # the DT baseline has no RTG head (cfgs/model/dt.yaml)
        spec += _mlp("decoder.predict_rtg", H, H, w.rtg_discretization * m.num_reward_components)
    if m.predict_future_states:
        spec += _mlp("decoder.predict_future_states", H, H, w.train_context_length * 2)
    return spec


def make_weights(cfg, seed: int = 0, head_gain: float = 4.0, still_bias: float = 0.0):
    """OrderedDict[str, np.ndarray(float32)].

    head_gain   scales the last linear of predict_action / predict_rtg so the categorical distributions are not
                near-uniform (xavier init gives logits of std ~0.3 over 1000 bins).
    still_bias  added to the action-head bias of bin 524 = discretised (accel 0, steer 0); a large
                value gives near-deterministic coasting, i.e. contact-free trajectories for simulator parity.
    """
    out = OrderedDict()
    for key, shape, kind in param_spec(cfg):
        n = int(np.prod(shape))
        u = uniform01(zlib.crc32(key.encode()) & 0x7FFFFFFF, n, seed)
        if kind in ("linear", "inproj", "seed"):
            fan_out, fan_in = (shape[0], shape[1]) if kind != "seed" else (shape[1], shape[2])
            if kind == "inproj":
                fan_out = fan_in  # reference treats the packed qkv as three square matrices (train_utils.py:36-40)
            bound = (6.0 / (fan_in + fan_out)) ** 0.5
            v = (2.0 * u - 1.0) * bound
        elif kind == "emb":
            v = (2.0 * u - 1.0) * (0.02 * 3.0 ** 0.5)  # uniform with std 0.02
        elif kind == "bias":
            v = (2.0 * u - 1.0) * 0.05
        elif kind == "ln_w":
            v = 1.0 + (2.0 * u - 1.0) * 0.1
        elif kind == "ln_b":
            v = (2.0 * u - 1.0) * 0.05
        else:
            raise ValueError(kind)
        out[key] = v.reshape(shape).astype(np.float32)
    for head in ("decoder.predict_action", "decoder.predict_rtg"):
        if f"{head}.mlp.3.weight" not in out:
            continue
        out[f"{head}.mlp.3.weight"] = (out[f"{head}.mlp.3.weight"] * np.float32(head_gain)).astype(np.float32)
    if still_bias:
        w = cfg.dataset.waymo
        still = int(np.round(0.5 * (w.accel_discretization - 1))) * w.steer_discretization + int(
            np.round(0.5 * (w.steer_discretization - 1)))
        out["decoder.predict_action.mlp.3.bias"][still] += np.float32(still_bias)
    return out
