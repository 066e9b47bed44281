"""ORACLE (test infrastructure only): CPU fp32 restatement of the CtRL-Sim policy network forward (M1-M10).

Follows, with plain tensor ops (linear / layer_norm / softmax), what the reference computes with stock torch modules:
  M1  get_causal_mask                       utils/train_utils.py:82-130
  M2  MapEncoder.forward                    modules/map_encoder.py:28-54
  M3  Encoder.forward embeddings            modules/encoder.py:50-153
  M4  nn.TransformerEncoder (post-LN, ReLU) modules/encoder.py:42-46,155-168
  M5-M7 nn.TransformerDecoder               modules/decoder.py:16-20,52
  M8/M9 heads                               modules/decoder.py:58,75 ; MLPLayer utils/layers.py:10-15
torch semantics restated (torch/nn/functional.py multi_head_attention_forward, nn.Transformer*Layer defaults):
packed in_proj (q,k,v), q scaled by d_head^-0.5, additive float mask, key-padding -> -inf, softmax, out_proj;
post-LN blocks x = LN(x + f(x)), eps 1e-5, ReLU FFN, no final stack norm.

It consumes the state-dict layout of the reference (ctrlsim_b200/weights.py) and the same `data` dict
(agent_states [B,A,T,8], agent_types [B,A,5], goals [B,A,5], actions [B,A,T], rtgs [B,A,T,3], timesteps [B,A,T,1],
road_points [B,P,Np,3], road_types [B,P,8]) and returns rtg_preds [B,A,T,1050], action_preds [B,A,T,1000].
Pinned against the reference modules by oracle/make_golden.py (--check-model) and tests/golden/model_*.npz.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def causal_mask_rule(A: int, T: int, K: int = 3, state_index: int = 0) -> torch.Tensor:
    """Boolean [L,L] 'allowed' matrix, token index = (t*A + a)*K + k with k in (state, rtg, action).

    Rule M1 (attend_own_return_action = False): j is visible from i iff t_j < t_i, or t_j == t_i and
    (k_j == state, or a_j == a_i and k_j <= k_i)."""
    L = A * T * K
    idx = torch.arange(L)
    t = idx // (A * K)
    a = (idx // K) % A
    k = idx % K
    ti, tj = t[:, None], t[None, :]
    same_agent = a[:, None] == a[None, :]
    # state_index: position of the state token inside a (step, agent) triple - 0 for CtRL-Sim (state, rtg, action),
    # 1 for the decision-transformer baseline (rtg, state, action) (utils/train_utils.py:85-88)
    allowed = (tj < ti) | ((tj == ti) & ((k[None, :] == state_index) | (same_agent & (k[None, :] <= k[:, None]))))
    return allowed


def _mlp(sd, prefix, x):
    x = F.linear(x, sd[f"{prefix}.mlp.0.weight"], sd[f"{prefix}.mlp.0.bias"])
    x = F.layer_norm(x, (x.shape[-1],), sd[f"{prefix}.mlp.1.weight"], sd[f"{prefix}.mlp.1.bias"], 1e-5)
    x = F.relu(x)
    return F.linear(x, sd[f"{prefix}.mlp.3.weight"], sd[f"{prefix}.mlp.3.bias"])


def _ln(sd, prefix, x):
    return F.layer_norm(x, (x.shape[-1],), sd[f"{prefix}.weight"], sd[f"{prefix}.bias"], 1e-5)


def _mha(sd, prefix, q_in, kv_in, n_heads, add_mask=None, key_pad=None):
    """q_in [B,Lq,H], kv_in [B,Lk,H]; add_mask [Lq,Lk] additive float; key_pad [B,Lk] bool True = ignore."""
    H = q_in.shape[-1]
    W, b = sd[f"{prefix}.in_proj_weight"], sd[f"{prefix}.in_proj_bias"]
    q = F.linear(q_in, W[:H], b[:H])
    k = F.linear(kv_in, W[H:2 * H], b[H:2 * H])
    v = F.linear(kv_in, W[2 * H:], b[2 * H:])
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    dh = H // n_heads
    q = q.view(B, Lq, n_heads, dh).transpose(1, 2) * (1.0 / math.sqrt(dh))
    k = k.view(B, Lk, n_heads, dh).transpose(1, 2)
    v = v.view(B, Lk, n_heads, dh).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if add_mask is not None:
        s = s + add_mask
    if key_pad is not None:
        s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Lq, H)
    return F.linear(o, sd[f"{prefix}.out_proj.weight"], sd[f"{prefix}.out_proj.bias"])


class ModelPort:
    def __init__(self, cfg, weights):
        self.cfg = cfg
        self.m, self.w = cfg.model, cfg.dataset.waymo
        self.sd = {k: torch.as_tensor(v, dtype=torch.float32) for k, v in weights.items()}
        A, T = self.w.max_num_agents, self.w.train_context_length
        self.dt = bool(getattr(self.m, "decision_transformer", False))
        allowed = causal_mask_rule(A, T, 3, state_index=1 if self.dt else 0)
        self.add_mask = torch.zeros(allowed.shape, dtype=torch.float32).masked_fill(~allowed, float("-inf"))

    def to(self, device):
        """Move the weights and the mask (bench.py gpu_torch_baseline: the same stock torch ops on a CUDA device)."""
        self.sd = {k: v.to(device) for k, v in self.sd.items()}
        self.add_mask = self.add_mask.to(device)
        return self

    # ---- M2 ------------------------------------------------------------------------------------------------
    def map_encoder(self, road_points, road_types):
        sd, H = self.sd, self.m.hidden_dim
        B, P, Np, _ = road_points.shape
        exist = road_points[..., -1]
        seg_invalid = exist.sum(dim=2) == 0
        pts_mask = (1.0 - exist).bool().view(B * P, Np).clone()
        pts_mask[:, 0][pts_mask.sum(-1) == Np] = False
        feats = _mlp(sd, "encoder.map_encoder.road_pts_encoder", road_points[..., :self.m.map_attr]).view(B * P, Np, H)
        seeds = sd["encoder.map_encoder.map_seeds"].view(1, 1, H).expand(B * P, 1, H)
        emb = _mha(sd, "encoder.map_encoder.road_pts_attn_layer", seeds, feats, self.m.num_heads, key_pad=pts_mask)
        emb = _ln(sd, "encoder.map_encoder.norm1", emb)
        emb2 = _ln(sd, "encoder.map_encoder.norm2", emb + _mlp(sd, "encoder.map_encoder.map_feats", emb))
        type_feats = _mlp(sd, "encoder.map_encoder.road_type_encoder", road_types).view(B * P, 1, H)
        out = _mlp(sd, "encoder.map_encoder.road_road_type_encoder", torch.cat([emb2, type_feats], dim=-1))
        return out.view(B, P, H), ~seg_invalid

    # ---- M3 + M4 -------------------------------------------------------------------------------------------
    def encoder(self, data):
        sd, H, A = self.sd, self.m.hidden_dim, self.w.max_num_agents
        st = data["agent_states"]
        B, _, T, _ = st.shape
        exist = st[..., -1:].transpose(1, 2)                                    # [B,T,A,1]
        types = data["agent_types"][:, None].expand(B, T, A, data["agent_types"].shape[-1])
        goals = data["goals"][:, None].expand(B, T, A, data["goals"].shape[-1])[..., :self.w.goal_dim].float()
        states = torch.cat([st[..., :-1].transpose(1, 2), types], dim=-1).float()    # [B,T,A,12]
        actions = data["actions"].transpose(1, 2).long()
        rtgs = data["rtgs"].transpose(1, 2).long()
        ts = data["timesteps"].transpose(1, 2)[..., 0].long()
        ids = torch.arange(A, device=st.device)[None, None, :].expand(B, T, A)
        ts_emb = sd["encoder.embed_timestep.weight"][ts]
        id_emb = sd["encoder.embed_agent_id.weight"][ids]
        s_emb = _mlp(sd, "encoder.embed_state", states)
        g_emb = _mlp(sd, "encoder.embed_goal", goals)
        s_emb = F.linear(torch.cat([s_emb, g_emb], dim=-1), sd["encoder.embed_state_goal.weight"],
                         sd["encoder.embed_state_goal.bias"]) + ts_emb + id_emb
        a_emb = sd["encoder.embed_action.weight"][actions] + ts_emb + id_emb
        if getattr(self.m, "decision_transformer", False):
            # DT baseline (cfgs/model/dt.yaml; modules/encoder.py:27-30,116-120): continuous, clip-normalised RTGs through
            # Linear(1 -> H) per component instead of embedding tables.  Oracle groundwork for SURVEY 8(f) N1.
            rc = data["rtgs"].transpose(1, 2).float()
            r_cat = torch.cat([F.linear(rc[..., k:k + 1], sd[f"encoder.embed_rtg_{nm}.weight"], sd[f"encoder.embed_rtg_{nm}.bias"])
                               for k, nm in enumerate(("goal", "veh", "road"))], dim=-1)
        else:
            r_cat = torch.cat([sd["encoder.embed_rtg_goal.weight"][rtgs[..., 0]],
                               sd["encoder.embed_rtg_veh.weight"][rtgs[..., 1]],
                               sd["encoder.embed_rtg_road.weight"][rtgs[..., 2]]], dim=-1)
        r_emb = F.linear(r_cat, sd["encoder.embed_rtg.weight"], sd["encoder.embed_rtg.bias"]) + ts_emb + id_emb
        ex = exist.float()
        s_emb, a_emb, r_emb = s_emb * ex, a_emb * ex, r_emb * ex
        init_emb = s_emb[:, 0]                                                   # [B,A,H] (pre-LN, encoder.py:111-112)
        init_exist = exist[:, 0, :, 0].bool()
        order = [r_emb, s_emb, a_emb] if self.dt else [s_emb, r_emb, a_emb]  # modules/encoder.py:139-152
        stacked = torch.stack(order, dim=3).reshape(B, T * A * 3, H)
        stacked = _ln(sd, "encoder.embed_ln", stacked)
        poly, valid = self.map_encoder(data["road_points"].float(), data["road_types"].float())
        x = torch.cat([poly, init_emb], dim=1)
        pad = ~torch.cat([valid, init_exist], dim=1)
        for l in range(self.m.num_transformer_encoder_layers):
            p = f"encoder.transformer_encoder.layers.{l}"
            x = _ln(sd, f"{p}.norm1", x + _mha(sd, f"{p}.self_attn", x, x, self.m.num_heads, key_pad=pad))
            ff = F.linear(F.relu(F.linear(x, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"])),
                          sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])
            x = _ln(sd, f"{p}.norm2", x + ff)
        return stacked, x, pad

    # ---- M5-M9 ---------------------------------------------------------------------------------------------
    def decoder(self, stacked, memory, pad):
        sd = self.sd
        x = stacked
        L = x.shape[1]
        mask = self.add_mask[:L, :L]
        for l in range(self.m.num_decoder_layers):
            p = f"decoder.transformer_decoder.layers.{l}"
            x = _ln(sd, f"{p}.norm1", x + _mha(sd, f"{p}.self_attn", x, x, self.m.num_heads, add_mask=mask))
            x = _ln(sd, f"{p}.norm2", x + _mha(sd, f"{p}.multihead_attn", x, memory, self.m.num_heads, key_pad=pad))
            ff = F.linear(F.relu(F.linear(x, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"])),
                          sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])
            x = _ln(sd, f"{p}.norm3", x + ff)
        return x

    @torch.no_grad()
    def forward(self, data):
        """data: dict of torch tensors (any float dtype; cast like the reference's .float()). Full-window forward."""
        A = self.w.max_num_agents
        stacked, memory, pad = self.encoder(data)
        B = stacked.shape[0]
        T = stacked.shape[1] // (A * 3)
        out = self.decoder(stacked, memory, pad).view(B, T * A, 3, -1)
        action_preds = _mlp(self.sd, "decoder.predict_action", out[:, :, 1]).view(B, T, A, -1).permute(0, 2, 1, 3)
        preds = {"action_preds": action_preds, "hidden": out}
        if getattr(self.m, "predict_rtg", True):  # the DT baseline has no RTG head (cfgs/model/dt.yaml)
            preds["rtg_preds"] = _mlp(self.sd, "decoder.predict_rtg", out[:, :, 0]).view(B, T, A, -1).permute(0, 2, 1, 3)
        return preds
