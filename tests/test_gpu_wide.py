"""GPU parity tests of the WIDE build (libctrlsim_b200_wide.so: 64 agents / 256 polylines per focal group, SURVEY 8(d)
config 2 "wide" variant): the kernels whose structure depends on the group geometry against dense torch references,
the forward pass against the oracle's network port and a closed-loop rollout against the oracle's rollout port, both
configured with the same caps (ctrlsim_b200.config.default_config(wide=True))."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

A, TOK_T = 64, 192


@pytest.fixture(scope="module")
def wlib():
    from ctrlsim_b200 import lib as L
    return L.load(wide=True)


@pytest.fixture(scope="module")
def wcfg():
    from ctrlsim_b200.config import default_config
    return default_config(wide=True)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(rc, lib):
    assert rc == 0, lib.ctrlsim_last_error()


def test_the_two_builds_reject_each_others_geometry(wcfg, dev):
    from ctrlsim_b200 import lib as L
    from ctrlsim_b200.config import default_config
    import ctypes as C
    for wide, cfg in ((False, wcfg), (True, default_config())):
        lib = L.load(wide=wide)
        h = C.c_void_p()
        cc = L.make_config(cfg)
        assert lib.ctrlsim_create(C.byref(cc), C.byref(h)) == -2
        assert b"specialised" in lib.ctrlsim_last_error()


@pytest.mark.parametrize("n_t", [1, 2, 5, 32])
def test_wide_attn_causal_matches_mask_rule(wlib, dev, n_t):
    """Rule M1 with 192 tokens per window step (3 x 64-key tiles per step) against the dense mask."""
    from oracle.model_port import causal_mask_rule
    G, L = 2, n_t * TOK_T
    g = torch.Generator(device="cpu").manual_seed(n_t)
    qkv = torch.randn(G, L, 768, generator=g).to(dev)
    O = torch.empty(G, L, 256, device=dev)
    _chk(wlib.ctrlsim_attn_causal(qkv.data_ptr(), O.data_ptr(), G, n_t, _stream()), wlib)
    allowed = causal_mask_rule(A, n_t, 3).to(dev)
    qh = qkv[..., :256].reshape(G, L, 8, 32).transpose(1, 2)
    kh = qkv[..., 256:512].reshape(G, L, 8, 32).transpose(1, 2)
    vh = qkv[..., 512:].reshape(G, L, 8, 32).transpose(1, 2)
    s = ((qh / math.sqrt(32)) @ kh.transpose(-1, -2)).masked_fill(~allowed, float("-inf"))
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(G, L, 256)
    assert (O - ref).abs().max().item() < 3e-5


@pytest.mark.parametrize("ti,own_row", [(0, True), (3, False), (31, True)])
def test_wide_attn_step_matches_mask_rule(wlib, dev, ti, own_row):
    from oracle.model_port import causal_mask_rule
    G, n_t = 2, ti + 1
    L = n_t * TOK_T
    g = torch.Generator(device="cpu").manual_seed(7 * ti + own_row)
    qkv = torch.randn(G, L, 768, generator=g).to(dev)
    rows = torch.randn(G, A, 768, generator=g).to(dev)
    O = torch.empty(G, A, 256, device=dev)
    _chk(wlib.ctrlsim_attn_step(qkv.data_ptr(), 768, 256, 512, L, rows.data_ptr(), O.data_ptr(), G, ti, 1 if own_row else 0,
                                _stream()), wlib)
    k_tok = 1 if own_row else 0
    idx = torch.tensor([(ti * A + a) * 3 + k_tok for a in range(A)], device=dev)
    full = qkv.clone()
    if own_row:
        full[:, idx, 256:] = rows[..., 256:]
    allowed = causal_mask_rule(A, n_t, 3).to(dev)[idx]
    qh = rows[..., :256].reshape(G, A, 8, 32).transpose(1, 2)
    kh = full[..., 256:512].reshape(G, L, 8, 32).transpose(1, 2)
    vh = full[..., 512:].reshape(G, L, 8, 32).transpose(1, 2)
    s = ((qh / math.sqrt(32)) @ kh.transpose(-1, -2)).masked_fill(~allowed, float("-inf"))
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(G, A, 256)
    assert (O - ref).abs().max().item() < 2e-5


def test_wide_attn_padded_320_keys_matches_torch(wlib, dev):
    """Attention over the 320 memory tokens (256 polylines + 64 initial-state tokens) with a key-padding mask."""
    G, Lq, Lk = 3, 300, 320
    g = torch.Generator(device="cpu").manual_seed(3)
    q = torch.randn(G, Lq, 256, generator=g).to(dev)
    kv = torch.randn(G, Lk, 512, generator=g).to(dev)
    pad = (torch.rand(G, Lk, generator=g) < 0.3).to(torch.uint8).to(dev)
    pad[:, 0] = 0
    O = torch.empty(G, Lq, 256, device=dev)
    _chk(wlib.ctrlsim_attn_padded(q.data_ptr(), 256, kv.data_ptr(), kv.data_ptr() + 256 * 4, 512, pad.data_ptr(), O.data_ptr(),
                                  G, Lq, Lk, _stream()), wlib)
    qh = q.reshape(G, Lq, 8, 32).transpose(1, 2)
    kh = kv[..., :256].reshape(G, Lk, 8, 32).transpose(1, 2)
    vh = kv[..., 256:].reshape(G, Lk, 8, 32).transpose(1, 2)
    s = ((qh / math.sqrt(32)) @ kh.transpose(-1, -2)).masked_fill(pad.bool()[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(G, Lq, 256)
    assert (O - ref).abs().max().item() < 2e-5


def test_wide_rollout_matches_oracle_port(wcfg, dev):
    """Closed loop with the wide caps: a dense 64-vehicle scene (focal groups of more than 24 members) and a scene whose
    288 polylines exceed the 256-polyline cap (the farthest are trimmed), first steps of the episode against the
    oracle's rollout port configured with the same caps: sampled RTG and action bins bit for bit, focal groups,
    positions."""
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.model import DeviceModel
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    from oracle.model_port import ModelPort
    from oracle.policy_port import RolloutPort
    weights = make_weights(wcfg, seed=4, still_bias=5.0)
    assert weights["encoder.embed_agent_id.weight"].shape == (64, 256)
    scenes = [make_scene(900, n_vehicles=64, n_roads=4, n_chunks=3), make_scene(901, n_vehicles=10, n_roads=12, n_chunks=3)]
    assert scenes[1]["preproc"]["road_points"].shape[0] > 256
    steps = 3
    pol = B200Policy(wcfg, "synthetic", DeviceModel(wcfg, weights, dev), seed=5, chunk_groups=64)
    ev = B200PolicyEvaluator(wcfg, pol, scenes=scenes)
    b = ev.build_batch(eval_threshold=64)
    ev.rollout(b, max_steps=steps)
    tr = b.trace()
    port = RolloutPort(wcfg, ModelPort(wcfg, weights), seed=5, eval_threshold=64)
    n_groups = 0
    for s, sc in enumerate(scenes):
        rec = port.run_scene(s, sc["json"], sc["preproc"], max_steps=steps)
        n = rec["n"]
        assert (tr["tr_rtg_idx"][s, :n, :steps].transpose(1, 0, 2) == rec["rtg_idx"][:steps]).all(), s
        assert (tr["tr_act_idx"][s, :n, :steps].T == rec["act_idx"][:steps]).all(), s
        assert np.abs(tr["tr_pos"][s, :n, :steps] - rec["pos"][:, :steps]).max() < 1e-3
        n_groups += len(rec["groups"][steps - 1])
        assert max(int((g["members"] >= 0).sum()) for g in rec["groups"][steps - 1]) > (24 if s == 0 else 0)  # wider than the default cap
    assert pol.groups_last_step == n_groups
