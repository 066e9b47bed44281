"""Device-side policy network handle: uploads a CtRL-Sim state dict and registers it with the CUDA library.

Accepts the reference's state-dict layout (keys ``encoder.*`` / ``decoder.*`` as saved by ``CtRLSim`` in
models/ctrl_sim.py:23-38; a Lightning checkpoint's ``state_dict`` works once its tensors are numpy/torch arrays).
A few tensors are folded on the host in float64 once at load time (they are constants of the network):

  derived.pool_U     [8,256]   U_h = W_k,h^T (q_h d_h^-0.5), q = map_seeds W_q^T + b_q   (modules/map_encoder.py:16-19,41)
  derived.pool_W/b   [256,2048] out_proj . blockdiag_h(W_v,h) and out_proj . b_v + out_proj.bias
  derived.pool_U2/W2/b2  the same three with road_pts_encoder.mlp.3 (W3, b3) folded in: U2_h = W3^T U_h, block h of W2 =
                     pool_W_h W3, b2 = pool_b + sum_h pool_W_h b3 - the pooling then runs on the hidden layer of the
                     point MLP and W3 is applied to 8 pooled vectors per polyline instead of 100 points (csrc/map_encoder.cu)
  derived.type_tab2  [9,256]   road_road_type_encoder.mlp.0[:, 256:] . road_type_encoder(one_hot_i) + bias; row 8 is
                               the all(-1) padding type (datasets/rl_waymo/dataset.py:425)
  derived.rtg_tab_*  [350,256] embed_rtg_{goal,veh,road}.weight folded through the matching block of embed_rtg
  derived.rtg_lin / rtg_lin_bias  [3,256] / [256]  decision-transformer variant only (cfgs/model/dt.yaml): the three
                     Linear(1, 256) RTG embeddings folded through embed_rtg (direction per component + one bias)
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib as _lib


def _np(v):
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy()
    return np.asarray(v)


def _mlp_f64(sd, prefix, x):
    h = x @ sd[f"{prefix}.mlp.0.weight"].T + sd[f"{prefix}.mlp.0.bias"]
    mu = h.mean(-1, keepdims=True)
    var = ((h - mu) ** 2).mean(-1, keepdims=True)
    h = (h - mu) / np.sqrt(var + 1e-5) * sd[f"{prefix}.mlp.1.weight"] + sd[f"{prefix}.mlp.1.bias"]
    h = np.maximum(h, 0.0)
    return h @ sd[f"{prefix}.mlp.3.weight"].T + sd[f"{prefix}.mlp.3.bias"]


def derive_weights(state_dict, cfg) -> dict:
    sd = {k: _np(v).astype(np.float64) for k, v in state_dict.items()}
    H, NH = cfg.model.hidden_dim, cfg.model.num_heads
    dh = H // NH
    me = "encoder.map_encoder"
    Win, bin_ = sd[f"{me}.road_pts_attn_layer.in_proj_weight"], sd[f"{me}.road_pts_attn_layer.in_proj_bias"]
    Wo, bo = sd[f"{me}.road_pts_attn_layer.out_proj.weight"], sd[f"{me}.road_pts_attn_layer.out_proj.bias"]
    q = (sd[f"{me}.map_seeds"].reshape(H) @ Win[:H].T + bin_[:H]) * (dh ** -0.5)
    Wk, Wv, bv = Win[H:2 * H], Win[2 * H:], bin_[2 * H:]
    U = np.stack([q[h * dh:(h + 1) * dh] @ Wk[h * dh:(h + 1) * dh] for h in range(NH)])
    pool_W = np.concatenate([Wo[:, h * dh:(h + 1) * dh] @ Wv[h * dh:(h + 1) * dh] for h in range(NH)], axis=1)
    pool_b = Wo @ bv + bo
    types_in = np.concatenate([np.eye(8), -np.ones((1, 8))], axis=0)
    tf = _mlp_f64(sd, f"{me}.road_type_encoder", types_in)
    Wrr, brr = sd[f"{me}.road_road_type_encoder.mlp.0.weight"], sd[f"{me}.road_road_type_encoder.mlp.0.bias"]
    tab2 = tf @ Wrr[:, H:].T + brr
    # second layer of road_pts_encoder folded through the (linear) pooling: csrc/map_encoder.cu
    W3, b3 = sd[f"{me}.road_pts_encoder.mlp.3.weight"], sd[f"{me}.road_pts_encoder.mlp.3.bias"]
    U2 = U @ W3
    pool_W2 = np.concatenate([pool_W[:, h * H:(h + 1) * H] @ W3 for h in range(NH)], axis=1)
    pool_b2 = pool_b + sum(pool_W[:, h * H:(h + 1) * H] @ b3 for h in range(NH))
    out = {"derived.pool_U": U, "derived.pool_W": pool_W, "derived.pool_b": pool_b, "derived.type_tab2": tab2,
           "derived.pool_U2": U2, "derived.pool_W2": pool_W2, "derived.pool_b2": pool_b2}
    Wr = sd["encoder.embed_rtg.weight"]
    if cfg.model.get("decision_transformer", False):
        # continuous RTG inputs (modules/encoder.py:27-30,116-120): embed_rtg([w_g r_g + b_g; w_v r_v + b_v; w_r r_r + b_r])
        # = r_g (Wr_g w_g) + r_v (Wr_v w_v) + r_r (Wr_r w_r) + (Wr_g b_g + Wr_v b_v + Wr_r b_r + b_rtg)
        names = ("goal", "veh", "road")
        out["derived.rtg_lin"] = np.stack([Wr[:, c * H:(c + 1) * H] @ sd[f"encoder.embed_rtg_{n}.weight"][:, 0]
                                           for c, n in enumerate(names)])
        out["derived.rtg_lin_bias"] = sd["encoder.embed_rtg.bias"] + sum(
            Wr[:, c * H:(c + 1) * H] @ sd[f"encoder.embed_rtg_{n}.bias"] for c, n in enumerate(names))
    else:
        for c, name in enumerate(("goal", "veh", "road")):
            out[f"derived.rtg_tab_{name}"] = sd[f"encoder.embed_rtg_{name}.weight"] @ Wr[:, c * H:(c + 1) * H].T
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}


class DeviceModel:
    """Owns the device copies of the weights and the library handle bound to them."""

    def __init__(self, cfg, state_dict, device="cuda:0"):
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CtrlSimError("ctrlsim_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.wide = _lib.is_wide(cfg)  # which build of the library serves this geometry (24/200 or 64/256)
        self.lib = _lib.load(wide=self.wide)
        torch.cuda.set_device(self.device)
        h = C.c_void_p()
        cc = _lib.make_config(cfg)
        _lib.check(self.lib.ctrlsim_create(C.byref(cc), C.byref(h)), "ctrlsim_create")
        self.handle = h
        self.tensors = {}
        skip = ("decoder.predict_future_states",)
        allw = {k: _np(v) for k, v in state_dict.items() if not k.startswith(skip)}
        allw.update(derive_weights(state_dict, cfg))
        for k, v in allw.items():
            t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(self.device)
            self.tensors[k] = t
            _lib.check(self.lib.ctrlsim_load_weights(self.handle, k.encode(), t.data_ptr(), t.numel()), f"load {k}")
        _lib.check(self.lib.ctrlsim_finalize_weights(self.handle), "ctrlsim_finalize_weights")
        self._ws = None
        self._ws_groups = 0

    @classmethod
    def load_from_checkpoint(cls, model_path, cfg, device="cuda:0", trusted=False):
        """Mirror of ``CtRLSim.load_from_checkpoint(model_path)`` (eval_sim.py:52): a Lightning ``.ckpt`` -> device model.
        Names and shapes are checked against ``cfg`` before anything is uploaded (checkpoint.load_checkpoint)."""
        from .checkpoint import load_checkpoint
        return cls(cfg, load_checkpoint(model_path, cfg, trusted=trusted), device)

    def workspace(self, groups: int):
        if self._ws is None or self._ws_groups < groups:
            nbytes = int(self.lib.ctrlsim_workspace_bytes(self.handle, groups))
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws_groups = groups
        return self._ws

    def eval(self):  # torch.nn.Module-like no-ops the reference Policy.__init__ calls (policies/policy.py:26-30)
        return self

    def forward_tokens(self, data: dict, n_t: int, rtg_idx_pass2):
        """Parity entry: data in the reference MotionData layout (numpy, one leading group dim). Returns
        (rtg_logits [G,24,1050], action_logits [G,24,1000]) as numpy."""
        dev = self.device
        G = data["agent_states"].shape[0]

        def f32(x):
            return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32), device=dev)

        def i32(x):
            return torch.as_tensor(np.ascontiguousarray(x, dtype=np.int32), device=dev)
        st, ty, go = f32(data["agent_states"]), f32(data["agent_types"]), f32(data["goals"])
        ac, rt = i32(data["actions"]), i32(data["rtgs"])
        ts = i32(np.asarray(data["timesteps"]).reshape(G, -1, 32)[:, 0] if np.asarray(data["timesteps"]).ndim > 2
                 else data["timesteps"])
        rp = f32(data["road_points"])
        rtypes = np.asarray(data["road_types"])
        if rtypes.ndim == 3:
            rtypes = np.where(rtypes.sum(-1) > 0, rtypes.argmax(-1), -1)
        rty = i32(rtypes)
        r2 = i32(rtg_idx_pass2)
        A = self.cfg.dataset.waymo.max_num_agents
        rtg_logits = torch.empty(G, A, 1050, dtype=torch.float32, device=dev)
        act_logits = torch.empty(G, A, 1000, dtype=torch.float32, device=dev)
        ws = self.workspace(G)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(self.lib.ctrlsim_forward_tokens(
            self.handle, G, n_t, n_t - 1, st.data_ptr(), ty.data_ptr(), go.data_ptr(), ac.data_ptr(), rt.data_ptr(),
            ts.data_ptr(), rp.data_ptr(), rty.data_ptr(), r2.data_ptr(), rtg_logits.data_ptr(), act_logits.data_ptr(),
            ws.data_ptr(), ws.numel(), stream), "ctrlsim_forward_tokens")
        torch.cuda.synchronize(dev)
        return rtg_logits.cpu().numpy(), act_logits.cpu().numpy()

    def forward_tokens_dt(self, data: dict, n_t: int):
        """Parity entry of the decision-transformer variant: data in the reference MotionData layout with continuous
        (clip-normalised) ``rtgs`` [G,A,32,3]. Returns action logits [G,A,1000] (state rows of window step n_t - 1)."""
        dev = self.device
        G = data["agent_states"].shape[0]

        def f32(x):
            return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32), device=dev)

        def i32(x):
            return torch.as_tensor(np.ascontiguousarray(x, dtype=np.int32), device=dev)
        st, ty, go = f32(data["agent_states"]), f32(data["agent_types"]), f32(data["goals"])
        ac, rt = i32(data["actions"]), f32(data["rtgs"])
        ts = i32(np.asarray(data["timesteps"]).reshape(G, -1, 32)[:, 0] if np.asarray(data["timesteps"]).ndim > 2
                 else data["timesteps"])
        rp = f32(data["road_points"])
        rtypes = np.asarray(data["road_types"])
        if rtypes.ndim == 3:
            rtypes = np.where(rtypes.sum(-1) > 0, rtypes.argmax(-1), -1)
        rty = i32(rtypes)
        A = self.cfg.dataset.waymo.max_num_agents
        act_logits = torch.empty(G, A, 1000, dtype=torch.float32, device=dev)
        ws = self.workspace(G)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(self.lib.ctrlsim_forward_tokens_dt(
            self.handle, G, n_t, n_t - 1, st.data_ptr(), ty.data_ptr(), go.data_ptr(), ac.data_ptr(), rt.data_ptr(),
            ts.data_ptr(), rp.data_ptr(), rty.data_ptr(), act_logits.data_ptr(), ws.data_ptr(), ws.numel(), stream),
            "ctrlsim_forward_tokens_dt")
        torch.cuda.synchronize(dev)
        return act_logits.cpu().numpy()

    def close(self):
        if self.handle:
            self.lib.ctrlsim_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
