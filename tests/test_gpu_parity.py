"""GPU parity tests (run on the B200 box: pytest -m gpu). Everything goes through the C-ABI of libctrlsim_b200.so.

Checkers: torch fp32 for the floating-point building blocks, the oracle (oracle/*.py, CPU) and the committed
reference fixtures (tests/golden/, produced by the unmodified reference - oracle/make_golden.py)."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from parity_checks import check_rollout_vs_reference

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-4   # |logit_gpu - logit_reference_fp32|, logits are O(1..10)
POS_TOL = 1e-3     # metres, BASELINE.json north_star


@pytest.fixture(scope="module")
def lib():
    from ctrlsim_b200 import lib as L
    return L.load()


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(rc, lib):
    assert rc == 0, lib.ctrlsim_last_error()


# ---------------------------------------------------------------------------------------------- building blocks
@pytest.mark.parametrize("M,N,K,relu", [(1, 256, 256, False), (200, 768, 256, True), (333, 1050, 256, False),
                                        (129, 256, 1024, False), (1000, 1000, 256, False), (77, 256, 2048, True)])
def test_linear_matches_torch(lib, dev, M, N, K, relu):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(dev)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    Cc = torch.empty(M, N, device=dev)
    _chk(lib.ctrlsim_linear(A.data_ptr(), W.data_ptr(), b.data_ptr(), Cc.data_ptr(), M, N, K, int(relu), _stream()), lib)
    ref = torch.nn.functional.linear(A.double(), W.double(), b.double())
    ref = torch.relu(ref) if relu else ref
    # tcgen05 accumulates with truncation: the error grows ~linearly with K/8 chained MMAs (gemm_tc.cu header)
    assert (Cc.double() - ref).abs().max().item() < 1e-5 * max(1.0, K / 256)


@pytest.mark.parametrize("name,M,relu", [("self_attn.in_proj_weight", 90 * 2304, False), ("linear1.weight", 70001, True),
                                         ("self_attn.out_proj.weight", 148 * 128 * 3 + 5, False), ("linear2.weight", 40000, False)])
def test_linear_with_registered_weights_matches_torch(cfg, lib, dev, name, M, relu):
    """Model weights are registered with the library (their lo tiles exist), which routes large K <= 256 layers to the
    weight-stationary kernel (gemm_tc_ws_kernel) and the rest to the streaming one: both against fp64, full output."""
    from ctrlsim_b200.model import DeviceModel
    from ctrlsim_b200.weights import make_weights
    model = DeviceModel(cfg, make_weights(cfg, seed=0), dev)
    W = model.tensors["decoder.transformer_decoder.layers.1." + name]
    N, K = W.shape
    g = torch.Generator(device="cpu").manual_seed(M % 1000)
    A = torch.randn(M, K, generator=g).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    C = torch.empty(M, N, device=dev)
    _chk(model.lib.ctrlsim_linear(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 1 if relu else 0, _stream()), model.lib)
    idx = torch.cat([torch.arange(0, 300, device=dev), torch.randint(0, M, (2000,), generator=g).to(dev), torch.arange(M - 300, M, device=dev)])
    ref = torch.nn.functional.linear(A[idx].double(), W.double(), b.double())
    if relu:
        ref = ref.relu()
    assert (C[idx].double() - ref).abs().max().item() < (2e-5 if K <= 256 else 5e-5)
    assert torch.isfinite(C).all()


@pytest.mark.parametrize("name,M", [("self_attn.out_proj.weight", 148 * 128 * 2 + 77), ("linear2.weight", 148 * 128 + 1),
                                    ("multihead_attn.out_proj.weight", 90 * 2304), ("linear2.weight", 300)])
def test_linear_res_ln_matches_torch(cfg, lib, dev, name, M):
    """X = LayerNorm(X + A W^T + b) in ONE kernel (gemm_tc_ta_ln_kernel: residual add and LayerNorm in the GEMM epilogue,
    a CTA owns whole rows) against torch in fp64, every row; K = 256 and K = 1024, a ragged last m-tile, and a small M
    that the fused kernel still serves (the model only sends it M >= 148 tiles)."""
    from ctrlsim_b200.model import DeviceModel
    from ctrlsim_b200.weights import make_weights
    model = DeviceModel(cfg, make_weights(cfg, seed=0), dev)
    pre = "decoder.transformer_decoder.layers.2."
    W = model.tensors[pre + name]
    gamma, beta = model.tensors[pre + "norm2.weight"], model.tensors[pre + "norm2.bias"]
    N, K = W.shape
    assert N == 256
    g = torch.Generator(device="cpu").manual_seed(M % 1000)
    A = torch.randn(M, K, generator=g).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    X = (torch.randn(M, N, generator=g) * 2.0 + 0.5).to(dev)
    gm = (gamma + 0.1 * torch.randn(N, generator=g).to(dev)).contiguous()
    bt = (beta + 0.1 * torch.randn(N, generator=g).to(dev)).contiguous()
    ref = torch.nn.functional.layer_norm(X.double() + torch.nn.functional.linear(A.double(), W.double(), b.double()), (N,),
                                         gm.double(), bt.double(), 1e-5)
    scratch = torch.empty(M, N, device=dev)
    Xc = X.clone()
    _chk(model.lib.ctrlsim_linear_res_ln(A.data_ptr(), W.data_ptr(), b.data_ptr(), Xc.data_ptr(), gm.data_ptr(), bt.data_ptr(),
                                         scratch.data_ptr(), M, K, _stream()), model.lib)
    err = (Xc.double() - ref).abs().max().item()
    assert err < (2e-5 if K <= 256 else 5e-5), err
    # the two-kernel path (what an unregistered W takes) gives the same to rounding
    W2 = W.clone()
    X2 = X.clone()
    _chk(model.lib.ctrlsim_linear_res_ln(A.data_ptr(), W2.data_ptr(), b.data_ptr(), X2.data_ptr(), gm.data_ptr(), bt.data_ptr(),
                                         scratch.data_ptr(), M, K, _stream()), model.lib)
    assert (X2.double() - ref).abs().max().item() < (2e-5 if K <= 256 else 5e-5)
    assert (X2 - Xc).abs().max().item() < 2e-5


@pytest.mark.parametrize("M,res,relu", [(1, False, False), (1000, True, False), (257, True, True)])
def test_layernorm_matches_torch(lib, dev, M, res, relu):
    g = torch.Generator(device="cpu").manual_seed(M)
    X = torch.randn(M, 256, generator=g).to(dev) * 3
    R = torch.randn(M, 256, generator=g).to(dev) if res else None
    gamma, beta = torch.randn(256, generator=g).to(dev), torch.randn(256, generator=g).to(dev)
    Y = torch.empty_like(X)
    _chk(lib.ctrlsim_layernorm(X.data_ptr(), R.data_ptr() if res else None, gamma.data_ptr(), beta.data_ptr(),
                               Y.data_ptr(), M, int(relu), _stream()), lib)
    ref = torch.nn.functional.layer_norm((X + R) if res else X, (256,), gamma, beta, 1e-5)
    ref = torch.relu(ref) if relu else ref
    assert (Y - ref).abs().max().item() < 2e-5


def test_attn_padded_matches_torch(lib, dev):
    G, Lq, Lk = 3, 150, 224
    g = torch.Generator(device="cpu").manual_seed(1)
    q = torch.randn(G, Lq, 256, generator=g).to(dev)
    kv = torch.randn(G, Lk, 512, generator=g).to(dev)
    pad = (torch.rand(G, Lk, generator=g) < 0.3).to(dev)
    pad[:, 0] = False
    O = torch.empty(G, Lq, 256, device=dev)
    padu8 = pad.to(torch.uint8).contiguous()
    _chk(lib.ctrlsim_attn_padded(q.data_ptr(), 256, kv.data_ptr(), kv.data_ptr() + 256 * 4, 512, padu8.data_ptr(),
                                 O.data_ptr(), G, Lq, Lk, _stream()), lib)
    qh = q.view(G, Lq, 8, 32).transpose(1, 2)
    kh = kv[..., :256].reshape(G, Lk, 8, 32).transpose(1, 2)
    vh = kv[..., 256:].reshape(G, Lk, 8, 32).transpose(1, 2)
    s = (qh / math.sqrt(32)) @ kh.transpose(-1, -2)
    s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(G, Lq, 256)
    assert (O - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("n_t", [1, 2, 7, 32])
def test_attn_causal_matches_mask_rule(lib, dev, n_t):
    """Rule M1 evaluated arithmetically == the dense additive mask of utils/train_utils.py:82-130 (pinned against the
    reference's get_causal_mask in tests/test_oracle.py)."""
    from oracle.model_port import causal_mask_rule
    G, L = 2, n_t * 72
    g = torch.Generator(device="cpu").manual_seed(n_t)
    qkv = torch.randn(G, L, 768, generator=g).to(dev)
    O = torch.empty(G, L, 256, device=dev)
    _chk(lib.ctrlsim_attn_causal(qkv.data_ptr(), O.data_ptr(), G, n_t, _stream()), lib)
    allowed = causal_mask_rule(24, n_t, 3).to(dev)
    qh = qkv[..., :256].reshape(G, L, 8, 32).transpose(1, 2)
    kh = qkv[..., 256:512].reshape(G, L, 8, 32).transpose(1, 2)
    vh = qkv[..., 512:].reshape(G, L, 8, 32).transpose(1, 2)
    s = ((qh / math.sqrt(32)) @ kh.transpose(-1, -2)).masked_fill(~allowed, float("-inf"))
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(G, L, 256)
    assert (O - ref).abs().max().item() < 3e-5


@pytest.mark.parametrize("ti,own_row,cache_layout", [(0, True, False), (0, False, False), (1, True, True), (5, False, False),
                                                      (31, True, False), (31, False, True)])
def test_attn_step_matches_mask_rule(lib, dev, ti, own_row, cache_layout):
    """The 24 rows of the current window step only (attention_step.cu): state rows of the last first-pass layer
    (own_row = False) and rtg rows of the second pass (own_row = True: the row's own NEW key replaces the first pass'
    rtg token) against the dense mask rule M1, for both K/V layouts (workspace QKV buffer / prefix-cache slot)."""
    from oracle.model_port import causal_mask_rule
    A, G, n_t = 24, 3, ti + 1
    L = n_t * 72
    g = torch.Generator(device="cpu").manual_seed(100 * ti + own_row)
    qkv = torch.randn(G, L, 768, generator=g).to(dev)          # first-pass rows
    rows = torch.randn(G, A, 768, generator=g).to(dev)         # recomputed rows of step ti (q | k | v)
    O = torch.empty(G, A, 256, device=dev)
    if cache_layout:  # [G, 2304, 512]: K | V, rows past the current length are stale
        group_rows = 2304
        kv = torch.randn(G, group_rows, 512, generator=g).to(dev)
        kv[:, :L] = qkv[..., 256:]
        args = (kv.data_ptr(), 512, 0, 256, group_rows)
    else:
        args = (qkv.data_ptr(), 768, 256, 512, L)
    _chk(lib.ctrlsim_attn_step(*args, rows.data_ptr(), O.data_ptr(), G, ti, 1 if own_row else 0, _stream()), lib)
    k_tok = 1 if own_row else 0
    idx = torch.tensor([(ti * A + a) * 3 + k_tok for a in range(A)], device=dev)
    full = qkv.clone()
    if own_row:  # the rows' own keys / values are the new ones
        full[:, idx, 256:] = rows[..., 256:]
    allowed = causal_mask_rule(A, n_t, 3).to(dev)[idx]           # [A, L]
    qh = rows[..., :256].reshape(G, A, 8, 32).transpose(1, 2)
    kh = full[..., 256:512].reshape(G, L, 8, 32).transpose(1, 2)
    vh = full[..., 512:].reshape(G, L, 8, 32).transpose(1, 2)
    s = ((qh / math.sqrt(32)) @ kh.transpose(-1, -2)).masked_fill(~allowed, float("-inf"))
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(G, A, 256)
    assert (O - ref).abs().max().item() < 2e-5


def test_map_pool_matches_torch(lib, dev):
    n_poly = 301  # more polylines than SMs: exercises the 2-stage TMA ring and the phase bookkeeping
    g = torch.Generator(device="cpu").manual_seed(5)
    feats = torch.randn(n_poly, 100, 256, generator=g).to(dev)
    U = (torch.randn(8, 256, generator=g) * 0.1).to(dev)
    valid = (torch.rand(n_poly, 100, generator=g) < 0.8)
    valid[3] = False          # all points masked -> point 0 is un-masked (map_encoder.py:31)
    valid[4, 1:] = False
    poly_valid = torch.ones(n_poly, dtype=torch.uint8)
    poly_valid[7] = 0
    pooled = torch.full((n_poly, 8, 256), float("nan"), device=dev)
    v8 = valid.to(torch.uint8).to(dev)
    pv = poly_valid.to(dev)
    _chk(lib.ctrlsim_map_pool(feats.data_ptr(), v8.data_ptr(), pv.data_ptr(), U.data_ptr(), pooled.data_ptr(), n_poly,
                              _stream()), lib)
    torch.cuda.synchronize()
    mask = ~valid.to(dev)
    mask[mask.all(-1), 0] = False
    s = torch.einsum("npd,hd->nhp", feats, U).masked_fill(mask[:, None, :], float("-inf"))
    ref = torch.einsum("nhp,npd->nhd", torch.softmax(s, -1), feats)
    ref[7] = 0
    assert torch.isfinite(pooled).all()
    assert (pooled - ref).abs().max().item() < 2e-5


def test_map_encode_pool_matches_the_unfused_chain(cfg, dev):
    """The fused polyline front end (map_encoder.cu: point MLP layer 1 -> scores -> masked softmax -> pooled HIDDEN
    vectors, W3 / value / output projections folded into derived.pool_W2) against the reference's chain in fp64:
    road_pts_encoder (both layers) -> scores against the folded seed query -> softmax with key padding (all-masked
    polylines un-mask point 0) -> weighted sum of the FEATURES -> value + output projection (modules/map_encoder.py:34-45)."""
    from ctrlsim_b200.model import DeviceModel, derive_weights
    from ctrlsim_b200.weights import make_weights
    weights = make_weights(cfg, seed=6)
    model = DeviceModel(cfg, weights, dev)
    lib = model.lib
    n_poly = 333
    g = torch.Generator(device="cpu").manual_seed(9)
    pts = torch.randn(n_poly, 100, 3, generator=g, dtype=torch.float64) * torch.tensor([40.0, 40.0, 0.0], dtype=torch.float64)
    valid = torch.rand(n_poly, 100, generator=g) > 0.2
    valid[5] = False                      # a padded polyline
    valid[7] = False; valid[7, 0] = True  # a single point
    pts[..., 2] = valid.double()
    pts[~valid] = 0.0
    poly_valid = valid.any(1)
    d_pts = pts.float().to(dev).contiguous()
    d_pv = poly_valid.to(torch.uint8).to(dev)
    pooled = torch.full((n_poly, 8, 256), float("nan"), device=dev)
    _chk(lib.ctrlsim_map_encode_pool(model.handle, d_pts.data_ptr(), d_pv.data_ptr(), pooled.data_ptr(), n_poly, _stream()), lib)
    sd = {k: torch.from_numpy(np.asarray(v)).double() for k, v in weights.items()}
    dv = {k: torch.from_numpy(v).double() for k, v in derive_weights(weights, cfg).items()}
    me = "encoder.map_encoder.road_pts_encoder"
    x = pts.float().double()
    h = torch.relu(torch.nn.functional.layer_norm(x @ sd[f"{me}.mlp.0.weight"].T + sd[f"{me}.mlp.0.bias"], (256,),
                                                  sd[f"{me}.mlp.1.weight"], sd[f"{me}.mlp.1.bias"], 1e-5))
    feats = h @ sd[f"{me}.mlp.3.weight"].T + sd[f"{me}.mlp.3.bias"]
    sc = feats @ dv["derived.pool_U"].T                                  # [n_poly, 100, 8]
    mask = valid.clone()
    mask[~poly_valid, 0] = True
    prob = torch.softmax(sc.masked_fill(~mask[..., None], float("-inf")), dim=1)
    pooled_feats = torch.einsum("nph,npd->nhd", prob, feats).reshape(n_poly, 8 * 256)
    want = pooled_feats @ dv["derived.pool_W"].T + dv["derived.pool_b"]  # what the encoder adds norm1 to
    got = pooled.double().cpu().reshape(n_poly, 8 * 256) @ dv["derived.pool_W2"].T + dv["derived.pool_b2"]
    ok = poly_valid
    assert torch.isfinite(pooled).all() and (pooled[~ok.to(dev)] == 0).all()
    assert (got[ok] - want[ok]).abs().max().item() < 2e-5 * max(1.0, want[ok].abs().max().item())


def test_sampler_bit_exact_vs_oracle(lib, dev):
    """Given identical fp32 inputs the device sampler returns exactly the oracle's index (integer CDF, explicit exp)."""
    from oracle import sampler
    rng = np.random.default_rng(0)
    for n, stride in ((1000, 1), (350, 3)):
        rows = 64
        x = (rng.standard_normal((rows, n * stride)) * rng.uniform(0.5, 8.0, (rows, 1))).astype(np.float32)
        ctr = rng.integers(0, 2 ** 31, (rows, 4)).astype(np.uint32)
        xd = torch.from_numpy(x).to(dev)
        cd = torch.from_numpy(ctr.view(np.int32)).to(dev)
        out = torch.empty(rows, dtype=torch.int32, device=dev)
        seed = 0x1234_5678_9ABC
        _chk(lib.ctrlsim_sample_rows(xd.data_ptr(), rows, n, n * stride, stride, seed, cd.data_ptr(), out.data_ptr(),
                                     _stream()), lib)
        got = out.cpu().numpy()
        want = np.array([sampler.sample_from_x(x[r, ::stride][:n], seed, *[int(c) for c in ctr[r]]) for r in range(rows)])
        assert (got == want).all(), (got, want)


def test_nucleus_sampler_bit_exact_vs_oracle(lib, dev):
    """Top-p filtering on the integer weights: the device's sort-free bisection == the oracle's sort-based definition,
    including heavy ties (quantised logits) and the edge thresholds."""
    from oracle import sampler
    rng = np.random.default_rng(1)
    rows, n = 96, 1000
    x = (rng.standard_normal((rows, n)) * rng.uniform(0.5, 8.0, (rows, 1))).astype(np.float32)
    x[rows // 2:] = np.round(x[rows // 2:] * 2) / 2  # many exact ties
    x[-1] = 0.0                                       # all categories tie
    ctr = rng.integers(0, 2 ** 31, (rows, 4)).astype(np.uint32)
    xd = torch.from_numpy(x).to(dev)
    cd = torch.from_numpy(ctr.view(np.int32)).to(dev)
    out = torch.empty(rows, dtype=torch.int32, device=dev)
    seed = 0x0BAD_5EED_1234
    for p in (0.0, 0.3, 0.8, 0.95, 1.0, 1.5):
        _chk(lib.ctrlsim_sample_rows_nucleus(xd.data_ptr(), rows, n, n, 1, seed, cd.data_ptr(), p, out.data_ptr(),
                                             _stream()), lib)
        got = out.cpu().numpy()
        want = np.array([sampler.sample_from_x_nucleus(x[r], p, seed, *[int(c) for c in ctr[r]]) for r in range(rows)])
        assert (got == want).all(), (p, np.nonzero(got != want)[0][:5], got[got != want][:5], want[got != want][:5])


def test_rollout_with_nucleus_sampling_matches_oracle_port(cfg, dev):
    """Closed loop with nucleus_sampling=True (cfgs/policy/ctrl_sim.yaml:10-11) == the oracle port, scene by scene."""
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    from ctrlsim_b200.model import DeviceModel
    from oracle.model_port import ModelPort
    from oracle.policy_port import RolloutPort
    weights = make_weights(cfg, seed=3, still_bias=2.0)
    scenes = [make_scene(60 + i, n_vehicles=6 + 2 * i, n_roads=2, n_chunks=3) for i in range(2)]
    steps = 4
    pol = B200Policy(cfg, "synthetic", DeviceModel(cfg, weights, dev), seed=9, nucleus_sampling=True, nucleus_threshold=0.8)
    ev = B200PolicyEvaluator(cfg, pol, scenes=scenes)
    b = ev.build_batch(eval_threshold=64)
    ev.rollout(b, max_steps=steps)
    tr = b.trace()
    port = RolloutPort(cfg, ModelPort(cfg, weights), seed=9, eval_threshold=64, nucleus=0.8)
    for s, sc in enumerate(scenes):
        rec = port.run_scene(s, sc["json"], sc["preproc"], max_steps=steps)
        n = rec["n"]
        assert (tr["tr_act_idx"][s, :n, :steps].T == rec["act_idx"][:steps]).all()
        assert (tr["tr_rtg_idx"][s, :n, :steps].transpose(1, 0, 2) == rec["rtg_idx"][:steps]).all()
        assert np.abs(tr["tr_pos"][s, :n, :steps] - rec["pos"][:, :steps]).max() < POS_TOL


def test_edge_scenes_match_oracle_port(cfg, dev):
    """Ragged / degenerate inputs in ONE batch: a scene without any road polyline (every focal is 'dead': action (0,0),
    autoregressive_policy.py:106-108,249-251), a single-vehicle scene, a scene whose vehicles all vanish early, and a
    64-vehicle scene (the grouping kernel's maximum) - against the oracle port, scene by scene."""
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene, preproc_from_json
    from ctrlsim_b200.weights import make_weights
    from ctrlsim_b200.model import DeviceModel
    from oracle.model_port import ModelPort
    from oracle.policy_port import RolloutPort
    weights = make_weights(cfg, seed=2, still_bias=5.0)
    no_roads = make_scene(80, n_vehicles=5, n_roads=1, n_chunks=2)
    no_roads["json"]["roads"] = []
    no_roads["preproc"] = preproc_from_json(no_roads["json"])
    single = make_scene(81, n_vehicles=1, n_roads=1, n_chunks=2)
    vanish = make_scene(82, n_vehicles=6, n_roads=1, n_chunks=3)
    for o in vanish["json"]["objects"]:  # every vehicle disappears after step 11
        for t in range(12, 91):
            o["position"][t] = {"x": -10000.0, "y": -10000.0}; o["velocity"][t] = {"x": -10000.0, "y": -10000.0}
            o["heading"][t] = -10000.0; o["valid"][t] = False
        o["goalPosition"] = dict(o["position"][11])
    full = make_scene(83, n_vehicles=64)
    scenes = [no_roads, single, vanish, full]
    steps = 14
    pol = B200Policy(cfg, "synthetic", DeviceModel(cfg, weights, dev), seed=4)
    ev = B200PolicyEvaluator(cfg, pol, scenes=scenes)
    b = ev.build_batch(eval_threshold=64)
    assert b.S == 4 and b.N == 64
    ev.rollout(b, max_steps=steps)
    tr = b.trace()
    port = RolloutPort(cfg, ModelPort(cfg, weights), seed=4, eval_threshold=64)
    for s, sc in enumerate(scenes):
        rec = port.run_scene(s, sc["json"], sc["preproc"], max_steps=steps if s < 3 else 2)
        n, T = rec["n"], (steps if s < 3 else 2)  # the 64-vehicle scene costs the CPU oracle ~24 forwards per step
        assert (tr["tr_act_idx"][s, :n, :T].T == rec["act_idx"][:T]).all(), s
        assert (tr["tr_rtg_idx"][s, :n, :T].transpose(1, 0, 2) == rec["rtg_idx"][:T]).all(), s
        assert (tr["tr_exist"][s, :n, :T] == rec["existence"][:, :T]).all(), s
        ex = rec["existence"][:, :T].astype(bool)
        assert np.abs(tr["tr_pos"][s, :n, :T] - rec["pos"][:, :T])[ex].max() < POS_TOL, s
        assert np.abs(tr["tr_action"][s, :n, :T] - np.stack([rec["accel"], rec["steer"]], -1)[:, :T])[ex].max() < 1e-4, s
    assert (tr["tr_act_idx"][0] == -1).all()  # no road polylines: nothing is ever sampled


def test_log_replay_batch_matches_oracle(cfg, dev):
    """BASELINE config 4 shape: every vehicle log-replayed (inverse bicycle -> FreeCar / Box2D integrate -> collision
    and off-road checks -> rewards) for whole 90-step episodes, a batch of scenes on the GPU vs the C simulator oracle
    scene by scene, for ALL 90 steps - through the vehicle-vehicle contacts several scenes run into (Box2D contact
    response on both sides; the simulator state is compared bit for bit)."""
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    from ctrlsim_b200.model import DeviceModel
    from oracle.policy_port import RolloutPort
    scenes = [make_scene(300 + i, n_vehicles=4 + (5 * i) % 29, n_roads=2 + i % 3, n_chunks=3 + i % 4, frac_short=0.25,
                         frac_parked=0.2 if i % 2 else 0.0) for i in range(40)]
    for sc in scenes[::5]:  # the first vehicle of some scenes leaves its lane at 20 degrees and crosses a road edge
        o = sc["json"]["objects"][0]
        th = math.radians(o["heading"][0]) + 0.35
        sp = max(6.0, math.hypot(o["velocity"][0]["x"], o["velocity"][0]["y"]))
        x0, y0 = o["position"][0]["x"], o["position"][0]["y"]
        for t, ok in enumerate(o["valid"]):
            if ok:
                o["position"][t] = {"x": x0 + sp * 0.1 * t * math.cos(th), "y": y0 + sp * 0.1 * t * math.sin(th)}
                o["heading"][t] = math.degrees(th)
                o["velocity"][t] = {"x": sp * math.cos(th), "y": sp * math.sin(th)}
        last = max(t for t, ok in enumerate(o["valid"]) if ok)
        o["goalPosition"] = dict(o["position"][last])
    pol = B200Policy(cfg, "synthetic", DeviceModel(cfg, make_weights(cfg, seed=0), dev), seed=0)
    ev = B200PolicyEvaluator(cfg, pol, scenes=scenes)
    b = ev.build_batch(eval_threshold=0, keep_replay_only=True)
    assert b.n_evaluated() == 0 and b.S == len(scenes)
    ev.rollout(b)
    tr = b.trace()
    port = RolloutPort(cfg, model=None, eval_threshold=0)
    n_veh_steps, n_coll_scenes, n_coll_steps, n_off = 0, 0, 0, 0
    T = 91
    for s, sc in enumerate(scenes):
        rec = port.run_scene(s, sc["json"], sc["preproc"], replay_only=True)
        n = rec["n"]
        ex = rec["existence"].astype(bool)
        assert (tr["tr_exist"][s, :n] == rec["existence"]).all()
        if not ex.any():
            continue
        # simulator state: bit-identical (same fp32 operation order, glibc trig), also while vehicles push each other
        assert (tr["tr_pos"][s, :n].astype(np.float64)[ex] == rec["pos"][ex]).all(), s
        assert (tr["tr_heading"][s, :n].astype(np.float64)[ex] == rec["heading"][ex]).all(), s
        assert np.abs(tr["tr_vel"][s, :n].astype(np.float64) - rec["vel"])[ex].max() < 1e-4
        assert np.abs(tr["tr_reward"][s, :n].astype(np.float64) - rec["reward"])[ex].max() < 1e-5  # incl. both collision flags
        assert np.abs(tr["tr_nearest"][s, :n, :, 0] - rec["nearest_dist"])[ex].max() < 1e-3
        exa = ex[:, :90]
        assert np.abs(tr["tr_action"][s, :n, :90] - np.stack([rec["accel"], rec["steer"]], -1)[:, :90])[exa].max() < 1e-4, s
        cv = (rec["reward"][:, :, 6] > 0) & ex
        n_veh_steps += int(ex.sum()); n_off += int((rec["reward"][:, :, 7][ex] > 0).sum())
        n_coll_scenes += bool(cv.any()); n_coll_steps += int(cv.sum())
    # the comparison is not vacuous: collisions that last (contact response at work) and off-road events all occur
    assert n_veh_steps > 30000 and n_off > 0 and n_coll_scenes >= 3 and n_coll_steps > 50, (n_veh_steps, n_off, n_coll_scenes, n_coll_steps)


def test_replay_workload_matches_the_cpu_reference_table(cfg, dev):
    """BASELINE config 4 (log replay of val_interactive-shaped scenes): the first 120 scenes of the workload of
    tools/replay_eval.py as one GPU batch against tests/golden/replay_oracle.json - the same scenes through the C
    restatement of the reference simulator (oracle/make_replay_table.py), 47 of them with vehicle-vehicle contacts.
    Collision / off-road vehicle-steps and vehicles are identical, the position and heading checksums bit-identical."""
    import json
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.model import DeviceModel
    from ctrlsim_b200.synth import make_replay_scene, replay_scene_summary
    from ctrlsim_b200.weights import make_weights
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "replay_oracle.json")))["rows"]
    scenes = [make_replay_scene(i) for i in range(len(ref))]
    pol = B200Policy(cfg, "synthetic", DeviceModel(cfg, make_weights(cfg, seed=0), dev), seed=0)
    ev = B200PolicyEvaluator(cfg, pol, scenes=scenes)
    b = ev.build_batch(eval_threshold=0, keep_replay_only=True)
    ev.rollout(b)
    tr = b.trace()
    gt = b.t["gt"].cpu().numpy()
    assert b.contact_overflow() == 0
    n_contact = 0
    for s, r in enumerate(ref):
        n = int(tr["n_veh"][s])
        got = replay_scene_summary(tr["tr_pos"][s, :n], tr["tr_heading"][s, :n], tr["tr_exist"][s, :n], tr["tr_reward"][s, :n], gt[s, :n, :, :2])
        for k in ("n", "veh_steps", "coll_steps", "off_steps", "coll_veh", "off_veh", "pos_sum", "head_sum"):
            assert got[k] == r[k], (s, k, got[k], r[k])
        assert abs(got["ade"] - r["ade"]) < 1e-12
        n_contact += r["coll_steps"] > 0
    assert n_contact >= 40


def test_geometry_known_answers(lib, dev):
    """Reference KATs: nocturne/cpp/tests/src/geometry/polygon_test.cc:60-86, intersection_test.cc:52-76."""
    eps = 1e-5

    def pp(a, b):
        A, B = torch.tensor(a, dtype=torch.float32, device=dev), torch.tensor(b, dtype=torch.float32, device=dev)
        o = torch.zeros(1, dtype=torch.int32, device=dev)
        _chk(lib.ctrlsim_geom_poly_poly(A.data_ptr(), len(a), B.data_ptr(), len(b), o.data_ptr(), _stream()), lib)
        return bool(o.item())

    def ps(a, s0, s1):
        A = torch.tensor(a, dtype=torch.float32, device=dev)
        Sg = torch.tensor([*s0, *s1], dtype=torch.float32, device=dev)
        o = torch.zeros(1, dtype=torch.int32, device=dev)
        _chk(lib.ctrlsim_geom_poly_seg(A.data_ptr(), len(a), Sg.data_ptr(), o.data_ptr(), _stream()), lib)
        return bool(o.item())
    sq = [(0, 0), (1, 0), (1, 1), (0, 1)]
    assert not pp(sq, [(1, 2), (2, 1), (2, 2)]) and not pp([(1, 2), (2, 1), (2, 2)], sq)
    assert pp(sq, [(1 - eps, 1 - eps), (2, 0), (2, 2)]) and pp([(1 - eps, 1 - eps), (2, 0), (2, 2)], sq)
    assert pp(sq, [(1, 1), (2, 0), (2, 2)]) and pp([(1, 1), (2, 0), (2, 2)], sq)   # touching vertex = intersects
    dia = [(1, 0), (0, 1), (-1, 0), (0, -1)]
    assert ps(dia, (0, 0.5), (0, -0.5)) and ps(dia, (-0.5, -0.5), (-0.5, -1.0))
    assert ps(dia, (-1, 0.5), (1, 1)) and ps(dia, (1, 1), (-1, -1))
    assert not ps(dia, (-1, -1 - eps), (1, -1 - eps)) and not ps(dia, (-3, 0.5), (-2, 1))


# ---------------------------------------------------------------------------------------------- network vs reference
def _model(cfg, spec, dev):
    from ctrlsim_b200.model import DeviceModel
    from ctrlsim_b200.weights import make_weights
    return DeviceModel(cfg, make_weights(cfg, **spec["weights"]), dev)


@pytest.mark.parametrize("name", ["plumbing", "crowded", "sparse"])
def test_forward_matches_reference_logits(cfg, dev, name):
    """Same tokens in -> RTG logits (pass 1) and action logits (pass 2, incremental) of the reference CtRLSim.forward."""
    g, spec, _ = load_golden(name)
    model = _model(cfg, spec, dev)
    for t in spec["logit_steps"]:
        n_t = min(t + 1, 32)
        ti = n_t - 1
        data = {k: g[f"in_{t}_{k}"][None] for k in ("agent_states", "agent_types", "goals", "actions", "road_points",
                                                    "road_types")}
        data["rtgs"] = g[f"in_{t}_rtgs_pass1"][None]
        data["timesteps"] = g[f"in_{t}_timesteps"][None][:, 0, :, 0]
        r2 = g[f"in_{t}_rtgs_pass2"][None][:, :, ti, :]
        rtg_logits, act_logits = model.forward_tokens(data, n_t, r2)
        n_real = int((g[f"in_{t}_agent_types"].sum(-1) > 0).sum())
        d_rtg = np.abs(rtg_logits[0, :n_real] - g[f"rtg_logits_{t}_0"][:n_real]).max()
        d_act = np.abs(act_logits[0, :n_real] - g[f"action_logits_{t}_0"][:n_real]).max()
        assert d_rtg < LOGIT_TOL and d_act < LOGIT_TOL, (name, t, d_rtg, d_act)


# ---------------------------------------------------------------------------------------------- closed loop vs reference
def _rollout(cfg, spec, dev, max_steps=None):
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    sc = make_scene(**spec["scene"])
    model = _model(cfg, spec, dev)
    tl = spec["tilts"]
    pol = B200Policy(cfg, "synthetic", model, tilt_dict={"tilt": True, "goal_tilt": tl[0], "veh_veh_tilt": tl[1],
                                                         "veh_edge_tilt": tl[2]}, seed=0)
    ev = B200PolicyEvaluator(cfg, pol, scenes=[sc])
    b = ev.build_batch(eval_threshold=64)
    ev.rollout(b, max_steps=max_steps)
    torch.cuda.synchronize()
    return ev, b, b.trace()


@pytest.mark.parametrize("name", ["plumbing", "crowded", "sparse"])
def test_rollout_matches_reference(cfg, dev, name):
    """Free-running closed loop vs the unmodified reference evaluator on the same scene JSON, weights and sampler seed,
    THROUGH vehicle-vehicle contacts (plumbing: from step 12, crowded: from step 16; Box2D contact response,
    sim_contacts.cuh).  Sampled bins are compared bit for bit up to the first marginal draw (see
    test_planner_adversary_matches_reference: a uniform within ~1e-5 of a CDF boundary, a few in 10^4 draws), which
    must lie well past the first contact and be an adjacent-bin flip of an RTG component; everything else is compared
    up to that step.  No such draw: whole episode and the final metrics."""
    g, spec, ref_metrics = load_golden(name)
    ev, b, tr = _rollout(cfg, spec, dev)
    t_r = check_rollout_vs_reference(tr, g, name)
    if t_r == 90:  # no marginal draw: the summary metrics must match the reference's compute_metrics
        m = ev.metrics_from_summary(ev.summarize(b))
        for k, v in ref_metrics.items():
            assert abs(m[k] - v) < 1e-4 * max(1.0, abs(v)), (k, m[k], v)


def test_config2_scene_rollout_matches_reference(cfg, dev):
    """ONE scene of BASELINE config 2's shape - 64 vehicles, all policy-controlled (~12 overlapping focal groups of 24
    per step), 256 polylines - against the unmodified reference evaluator for 44 steps: 12 of them in the sliding-window
    phase (t >= 32), vehicle-vehicle contacts from step 14 on (149 vehicle-steps in collision).  Sampled bins bit for
    bit (11,000 draws; a marginal draw, if any, must come after the window started sliding), states bit for bit."""
    g, spec, ref_metrics = load_golden("config2")
    c = cfg.copy()
    c.nocturne = cfg.nocturne.copy()
    c.nocturne.steps = spec["steps"]
    ev, b, tr = _rollout(c, spec, dev)
    n = g["pos"].shape[0]
    assert n == 64 and b.Pm == 256 and len(g["evaluated"]) == 64
    t_r = check_rollout_vs_reference(tr, g, "config2", steps=spec["steps"], min_flip=34)
    if t_r == spec["steps"]:
        m = ev.metrics_from_summary(ev.summarize(b))
        for k, v in ref_metrics.items():
            assert abs(m[k] - v) < 1e-4 * max(1.0, abs(v)), (k, m[k], v)


@pytest.mark.parametrize("name", ["plumbing", "crowded", "sparse"])
def test_simulator_with_contacts_reproduces_reference_episodes(cfg, dev, name):
    """Simulator only, no sampling: the controls the reference evaluator applied to the policy-controlled vehicles are
    fed to ctrlsim_sim_step step by step (log-replayed vehicles compute their own inverse-bicycle controls); the
    trajectories of ALL 90 steps must be the reference's - through the contacts of 'plumbing' (from step 12) and
    'crowded' (30 vehicles, from step 16), where a contact-free simulator is metres off."""
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    g, spec, _ = load_golden(name)
    sc = make_scene(**spec["scene"])
    pol = B200Policy(cfg, "synthetic", _model(cfg, spec, dev), seed=0)
    ev = B200PolicyEvaluator(cfg, pol, scenes=[sc])
    b = ev.build_batch(eval_threshold=64)
    assert b.evaluated_ids[0] == sorted(int(v) for v in g["evaluated"])
    pol.reset(b)
    n = g["pos"].shape[0]
    ctrl = torch.from_numpy(np.stack([g["accel"], g["steer"]], -1)).to(dev)  # [n, 91, 2]
    for t in range(90):
        pol.update_state(b, t)
        b.t["next_action"][0, :n] = ctrl[:, t]
        pol.act(b, t)
    pol.update_state(b, 90)
    tr = b.trace()
    ex = g["existence"].astype(bool)
    assert (tr["tr_exist"][0, :n] == g["existence"]).all()
    dpos = np.abs(tr["tr_pos"][0, :n].astype(np.float64) - g["pos"])[ex].max()
    dhead = np.abs(tr["tr_heading"][0, :n].astype(np.float64) - g["heading"])[ex].max()
    coll = (tr["tr_reward"][0, :n, :, 6] == 1)[ex]
    # the simulator replays the reference's fp32 operation order with glibc's own sinf / cosf / tanf (glibc_trig.h, the
    # default since round 2): bit-identical through the contacts of plumbing and crowded (30 vehicles, 74 steps in contact)
    assert dpos == 0.0 and dhead == 0.0, (name, dpos, dhead)
    assert (coll == (g["reward"][:, :, 6] == 1)[ex]).all()
    if name != "sparse":
        assert coll.any()


def test_rollout_matches_oracle_port_multi_scene(cfg, dev):
    """Batched rollout of several scenes == oracle port run scene by scene (first 6 steps; covers batching/compaction)."""
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    from ctrlsim_b200.model import DeviceModel
    from oracle.model_port import ModelPort
    from oracle.policy_port import RolloutPort
    weights = make_weights(cfg, seed=3, still_bias=6.0)
    scenes = [make_scene(20 + i, n_vehicles=5 + 3 * i, n_roads=2, n_chunks=3) for i in range(3)]
    steps = 6
    pol = B200Policy(cfg, "synthetic", DeviceModel(cfg, weights, dev), seed=7)
    ev = B200PolicyEvaluator(cfg, pol, scenes=scenes)
    b = ev.build_batch(eval_threshold=64)
    ev.rollout(b, max_steps=steps)
    tr = b.trace()
    port = RolloutPort(cfg, ModelPort(cfg, weights), seed=7, eval_threshold=64)
    for s, sc in enumerate(scenes):
        rec = port.run_scene(s, sc["json"], sc["preproc"], max_steps=steps)
        n = rec["n"]
        assert (tr["tr_act_idx"][s, :n, :steps].T == rec["act_idx"][:steps]).all()
        assert (tr["tr_rtg_idx"][s, :n, :steps].transpose(1, 0, 2) == rec["rtg_idx"][:steps]).all()
        assert np.abs(tr["tr_pos"][s, :n, :steps] - rec["pos"][:, :steps]).max() < POS_TOL


@pytest.mark.gpu
def test_caches_do_not_change_results(cfg, dev, monkeypatch):
    """Steps 0..31 with the polyline-encoder cache and the decoder prefix cache (incremental decode) == the same steps
    with every group recomputed from scratch: identical sampled bins, trajectories within float tolerance; and the
    incremental path is really taken.  Multi-chunk (chunk_groups smaller than the batch) on purpose."""
    import ctypes as C
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    from ctrlsim_b200.model import DeviceModel
    weights = make_weights(cfg, seed=5, still_bias=5.0)
    scenes = [make_scene(40 + i, n_vehicles=20 + 2 * i, n_roads=3, n_chunks=4, frac_short=0.3) for i in range(12)]
    steps = 36  # crosses t = 32, where both caches stop applying
    traces = {}
    for mode in ("cached", "plain"):
        monkeypatch.setenv("CTRLSIM_MAP_CACHE", "1" if mode == "cached" else "0")
        monkeypatch.setenv("CTRLSIM_PREFIX_CACHE", "1" if mode == "cached" else "0")
        model = DeviceModel(cfg, weights, dev)
        pol = B200Policy(cfg, "synthetic", model, seed=11, chunk_groups=48)
        ev = B200PolicyEvaluator(cfg, pol, scenes=scenes)
        b = ev.build_batch(eval_threshold=64)
        ev.rollout(b, max_steps=steps)
        traces[mode] = b.trace()
        if mode == "cached":
            inc, full = C.c_int64(0), C.c_int64(0)
            pol.lib.ctrlsim_prefix_cache_stats(model.handle, C.byref(inc), C.byref(full))
            assert inc.value >= 31 and full.value >= 1, (inc.value, full.value)
            hit, miss = C.c_int64(0), C.c_int64(0)
            pol.lib.ctrlsim_map_cache_stats(model.handle, C.byref(hit), C.byref(miss))
            assert miss.value >= 1
    a, p = traces["cached"], traces["plain"]
    assert (a["tr_rtg_idx"][:, :, :steps] == p["tr_rtg_idx"][:, :, :steps]).all()
    assert (a["tr_act_idx"][:, :, :steps] == p["tr_act_idx"][:, :, :steps]).all()
    assert (a["tr_exist"] == p["tr_exist"]).all()
    assert np.abs(a["tr_pos"][:, :, :steps].astype(np.float64) - p["tr_pos"][:, :, :steps]).max() < POS_TOL


@pytest.mark.gpu
def test_scene_results_do_not_depend_on_batch_composition(cfg, dev):
    """Size-independent property at BASELINE config-2 scene size (64 vehicles x 256 polylines, several focal groups
    per scene): a scene's rollout is a function of (scene, global scene id, seed) only - not of which other scenes
    share the batch, their order, or how focal groups fall into chunks. This is what makes sharding scenes over GPUs
     and chunking exact: sampler counters are keyed by global scene id / vehicle / step, every kernel
    computes a row or a group from that row's or group's data alone. Compared bit for bit."""
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    from ctrlsim_b200.model import DeviceModel
    weights = make_weights(cfg, seed=2)
    ids = [300 + i for i in range(6)]
    scenes = [make_scene(i) for i in ids]  # default = config-2 shape
    steps = 12
    model = DeviceModel(cfg, weights, dev)

    def run(order, chunk):
        pol = B200Policy(cfg, "synthetic", model, seed=5, chunk_groups=chunk)
        ev = B200PolicyEvaluator(cfg, pol, scenes=[scenes[k] for k in order], scene_ids=[ids[k] for k in order])
        b = ev.build_batch(eval_threshold=64)
        ev.rollout(b, max_steps=steps)
        tr = b.trace()
        return {ids[k]: {f: tr[f][j] for f in ("tr_act_idx", "tr_rtg_idx", "tr_pos", "tr_exist")}
                for j, k in enumerate(order)}, pol.groups_last_step

    base, groups = run(list(range(6)), 256)
    assert groups > 6  # several focal groups per scene
    for order, chunk in (([5, 4, 3, 2, 1, 0], 7), ([2, 0, 5], 256), ([4], 3)):
        other, _ = run(order, chunk)
        for sid, rec in other.items():
            for f, v in rec.items():
                assert np.array_equal(v, base[sid][f]), (sid, f, order, chunk)


def test_full_size_batch_is_deterministic_and_shard_invariant(cfg, dev):
    """BASELINE config 2 at its FULL size - 256 scenes x 64 vehicles x 256 polylines, 16,384 controlled agents, ~2,900
    focal groups per step - for 34 steps (the whole cached phase and the first two sliding-window steps), through
    properties that need no oracle: (1) the run is reproducible bit for bit with a different chunking of the focal
    groups (prefix-cache slots, workspace chunks and GEMM row counts all change); (2) four 64-scene shards, as four
    ranks would own them (scene k -> rank k mod 4, global scene ids kept), reproduce every scene's draws and
    trajectories bit for bit, and the sum of their metric summaries equals the full batch's summary - the quantity
    the evaluation's one all-reduce adds up."""
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    from ctrlsim_b200.model import DeviceModel
    S, steps = 256, 34
    ids = list(range(1000, 1000 + S))
    scenes = [make_scene(i) for i in ids]
    model = DeviceModel(cfg, make_weights(cfg, seed=0), dev)
    fields = ("tr_act_idx", "tr_rtg_idx", "tr_pos", "tr_heading", "tr_exist", "tr_reward")

    def run(sel, chunk):
        pol = B200Policy(cfg, "synthetic", model, seed=0, chunk_groups=chunk)
        ev = B200PolicyEvaluator(cfg, pol, scenes=[scenes[k] for k in sel], scene_ids=[ids[k] for k in sel])
        b = ev.build_batch(eval_threshold=64)
        ev.rollout(b, max_steps=steps)
        tr = b.trace()
        summ = ev.summarize(b)
        return {f: tr[f] for f in fields}, tr["n_veh"], summ, pol.groups_last_step, b.n_evaluated()

    full, n_veh, summ, groups, n_eval = run(list(range(S)), 256)
    assert n_eval == S * 64 and groups > 2500
    assert (full["tr_act_idx"][:, :, 9:steps] >= 0).any() and np.isfinite(full["tr_pos"][:, :, :steps + 1]).all()
    again, _, summ2, _, _ = run(list(range(S)), 96)
    for f in fields:
        assert np.array_equal(full[f], again[f]), ("chunking changed", f)
    assert np.array_equal(summ, summ2)
    total = np.zeros_like(summ)
    for r in range(4):
        sel = list(range(r, S, 4))
        part, _, sm, _, _ = run(sel, 256)
        for f in fields:
            assert np.array_equal(part[f], full[f][sel]), ("sharding changed", f, r)
        total += sm
    # integer counts and histograms add up exactly; the float sums (ADE / FDE, per-scene rates) to rounding
    assert np.allclose(total, summ, rtol=1e-12, atol=1e-9) and np.array_equal(total[8:], summ[8:])


# ---------------------------------------------------------------------------------------------- planner vs adversary
@pytest.mark.parametrize("name", ["policies", "cat"])
def test_planner_adversary_matches_reference(cfg, dev, name):
    """ego driven by the planner policy, one other vehicle by the adversary (a second policy with its
    own tilts and sampler seed, or the scripted CAT trajectory), everyone else log-replayed - the whole 90-step episode
    of every scene at once vs the unmodified reference PlannerAdversaryEvaluator (tests/golden/planner_adversary_*.npz,
    contact-free scenes).

    Sampled bins are compared bit for bit.  Two fp32-class implementations agree on a categorical draw unless the
    uniform lands within their logit difference (here ~7e-6) of a CDF boundary; with ~60 effective RTG bins that is
    a few draws in 10^4, and these fixtures sample ~3000.  Such a marginal draw may move an RTG sample to the
    ADJACENT bin; from there on the two rollouts legitimately differ (a different RTG token is in the context).  So:
    every draw before a policy's first marginal flip must be identical, the flip must come after the sliding-window
    regime has started (t >= 32: log replay, hand-over at t = 9, cached and full-window steps are all covered bit-exact),
    it must be an adjacent-bin flip of an RTG component, and trajectories are compared up to it."""
    from conftest import load_planner_adversary_golden
    from ctrlsim_b200.evaluator import B200Policy
    from ctrlsim_b200.model import DeviceModel
    from ctrlsim_b200.planner_adversary import B200PlannerAdversaryEvaluator, CatAdversary, vehicle_index_of_object
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    recs, spec, ref_metrics = load_planner_adversary_golden(name)
    scenes = [make_scene(**s) for s in spec["scenes"]]
    weights = make_weights(cfg, **spec["weights"])
    tp, ta = spec["tilts_planner"], spec["tilts_adversary"]
    tilt = lambda t: {"tilt": True, "goal_tilt": t[0], "veh_veh_tilt": t[1], "veh_edge_tilt": t[2]}
    planner = B200Policy(cfg, "synthetic", DeviceModel(cfg, weights, dev), tilt_dict=tilt(tp), seed=spec["seeds"][0])
    if spec["cat"]:
        adversary = CatAdversary()
    else:
        adversary = B200Policy(cfg, "synthetic", DeviceModel(cfg, weights, dev), tilt_dict=tilt(ta), seed=spec["seeds"][1])
    pairs = [tuple(vehicle_index_of_object(sc["json"], o) for o in p) for sc, p in zip(scenes, spec["pairs"])]
    ev = B200PlannerAdversaryEvaluator(cfg, planner, adversary, scenes=scenes, pairs=pairs,
                                       adv_trajs=[r["adv_pos"] for r in recs])
    metrics, _ = ev.evaluate_planner_adversary()
    tr = ev.batch.trace()
    T = 90
    views = {"planner": ev.view_planner, "adversary": ev.view_adversary}
    exact = True
    for s, g in enumerate(recs):
        assert not (g["reward"][:, :, 6] * g["existence"]).any(), "fixture must be contact-free"
        n = g["pos"].shape[0]
        assert tuple(int(x) for x in g["ego_adv"]) == pairs[s]
        assert (tr["tr_exist"][s, :n, :T + 1] == g["existence"]).all()
        t_ok = T  # steps [0, t_ok) are free of marginal flips in every policy of this scene
        for role, view in views.items():
            if view is None:
                continue
            act = view.t["tr_act_idx"][s, :n, :T].cpu().numpy().T
            rtg = view.t["tr_rtg_idx"][s, :n, :T].cpu().numpy().transpose(1, 0, 2).astype(np.int64)
            bad_r = np.argwhere(rtg != g[f"{role}_rtg_idx"])
            bad_a = np.argwhere(act != g[f"{role}_act_idx"])
            t_r = int(bad_r[:, 0].min()) if len(bad_r) else T
            t_a = int(bad_a[:, 0].min()) if len(bad_a) else T
            assert t_a > t_r or t_a == T, (role, "an action draw differs before any RTG draw did", t_a, t_r)
            if t_r < T:
                exact = False
                assert t_r >= 32, (role, s, "draws differ before the sliding-window regime", bad_r[:4].tolist())
                first = bad_r[bad_r[:, 0] == t_r]
                for t, v, c in first:
                    assert abs(int(rtg[t, v, c]) - int(g[f"{role}_rtg_idx"][t, v, c])) == 1, (role, t, v, c)
            t_ok = min(t_ok, t_r)
        ex = g["existence"][:, :t_ok + 1].astype(bool)
        dpos = np.abs(tr["tr_pos"][s, :n, :t_ok + 1].astype(np.float64) - g["pos"][:, :t_ok + 1])[ex].max()
        ga = np.stack([g["accel"], g["steer"]], -1)
        dacc = np.abs(tr["tr_action"][s, :n, :t_ok] - ga[:, :t_ok])[ex[:, :t_ok]].max()  # controls applied before the flip
        assert dpos < POS_TOL and dacc < 2e-4, (s, t_ok, dpos, dacc)
    for k, v in ref_metrics.items():
        if isinstance(v, float) and math.isnan(v):
            assert math.isnan(metrics[k]), (k, metrics[k])
        elif exact:
            assert abs(metrics[k] - v) < 1e-4 * max(1.0, abs(v)), (k, metrics[k], v)
        else:  # a marginal flip happened late in some episode: the aggregate metrics can only move a little
            assert abs(metrics[k] - v) < 0.05 * max(1.0, abs(v)), (k, metrics[k], v)


def test_fp64_trig_mode_stays_within_tolerance_through_contacts(cfg, dev, monkeypatch):
    """CTRLSIM_TRIG=fp64 (sinf / cosf evaluated in fp64 and rounded once - the default of round 1) differs from glibc's
    not-always-correctly-rounded sinf / cosf by 1 ulp in ~1 % of calls; 'crowded' (30 vehicles pushing each other from
    step 16 on) replayed with the reference's controls then drifts by 0.73 mm / 2.3e-5 rad - inside the 1e-3 m bar, and
    exactly what the CPU emulation of that arithmetic predicts (test_gpu_acceptance_logic_on_cpu_emulation_of_gpu_arithmetic).
    The default mode (glibc's algorithm) is bit-exact: test_simulator_with_contacts_reproduces_reference_episodes."""
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    g, spec, _ = load_golden("crowded")
    monkeypatch.setenv("CTRLSIM_TRIG", "fp64")
    try:
        model = _model(cfg, spec, dev)  # ctrlsim_create reads the switch
        pol = B200Policy(cfg, "synthetic", model, seed=0)
        ev = B200PolicyEvaluator(cfg, pol, scenes=[make_scene(**spec["scene"])])
        b = ev.build_batch(eval_threshold=64)
        pol.reset(b)
        n = g["pos"].shape[0]
        ctrl = torch.from_numpy(np.stack([g["accel"], g["steer"]], -1)).to(dev)
        for t in range(90):
            pol.update_state(b, t)
            b.t["next_action"][0, :n] = ctrl[:, t]
            pol.act(b, t)
        pol.update_state(b, 90)
        tr = b.trace()
    finally:
        monkeypatch.delenv("CTRLSIM_TRIG")
        _model(cfg, spec, dev)  # the switch is process-wide: the next handle sets it back to the default
    ex = g["existence"].astype(bool)
    dpos = np.abs(tr["tr_pos"][0, :n].astype(np.float64) - g["pos"])[ex].max()
    dhead = np.abs(tr["tr_heading"][0, :n].astype(np.float64) - g["heading"])[ex].max()
    assert abs(dpos - 0.000732421875) < 1e-9 and abs(dhead - 2.3126602172851562e-05) < 1e-9, (dpos, dhead)


def test_reference_signature_adapter_matches_the_batched_path(cfg, dev):
    """SURVEY 8(b), GPU half: ``B200AutoregressivePolicy`` driven through the reference's own call sequence -
    reset(vehicle_data_dict), then per step update_state(vdd, v2e, t) / predict(vdd, gt, preproc, dset, v2e, t) / act(veh,
    t, vdd) with a dict-of-lists world state and a CPU simulator owned by the caller (here the oracle's simulator and
    evaluator-loop restatement stand in for nocturne_cpp and PolicyEvaluator, which do not exist on the GPU box; the
    stock PolicyEvaluator itself drives the same class in tests/test_oracle.py) - must draw the bins the unmodified
    reference drew on the 'sparse' fixture, across the switch to the sliding window at t = 32."""
    from ctrlsim_b200.policy_adapter import B200AutoregressivePolicy
    from ctrlsim_b200.synth import make_scene
    from oracle.policy_port import RolloutPort
    g, spec, _ = load_golden("sparse")
    sc = make_scene(**spec["scene"])
    T = 36
    kd = {"next_acceleration": "next_acceleration", "next_steering": "next_steering", "rtgs": "rtgs"}
    tl = spec.get("tilts", [0, 0, 0])
    td = {"tilt": True, "goal_tilt": tl[0], "veh_veh_tilt": tl[1], "veh_edge_tilt": tl[2]}
    policy = B200AutoregressivePolicy(cfg, "synthetic", _model(cfg, spec, dev), True, True, True, False, False, False, False, kd,
                                      td, "ctrl_sim", 1.0, False, 0.8, seed=0)
    port = RolloutPort(cfg, None)
    ctx = port.setup_scene(0, sc["json"])
    n, sim, rec, gt = ctx["n"], ctx["sim"], ctx["rec"], ctx["gt"]
    v2e = [int(v) for v in g["evaluated"]]
    gt_data_dict = {v: {"traj": [list(gt[v, t]) for t in range(91)]} for v in range(n)}
    vdd = {v: {"position": [], "velocity": [], "heading": [], "existence": [], "acceleration": [], "steering": [], "timestep": [],
               "rtgs": [], "goal_position": {"x": ctx["goal"][v, 0], "y": ctx["goal"][v, 1]}, "goal_heading": ctx["goal"][v, 2],
               "goal_speed": ctx["goal"][v, 3], "length": float(ctx["parsed"]["size"][v, 0]), "width": float(ctx["parsed"]["size"][v, 1]),
               "type": "vehicle", "next_acceleration": 0.0, "next_steering": 0.0} for v in range(n)}

    class Veh:  # the slice of the pybind Vehicle that act() touches
        def __init__(self, v): self.v, self.acceleration, self.steering, self.braked, self.parked = v, None, None, None, False
        def getID(self): return self.v
        def brake(self, x): self.braked = x
        def setPosition(self, x, y): self.parked = True

    policy.scene_index = -1  # reset() advances it to 0 = the fixture's scene id
    policy.reset(vdd)
    next_act = np.zeros((n, 2))
    for t in range(T):
        port.observe(ctx, t)
        for v in range(n):
            d = vdd[v]
            d["position"].append({"x": rec["pos"][v, t, 0], "y": rec["pos"][v, t, 1]})
            d["velocity"].append({"x": rec["vel"][v, t, 0], "y": rec["vel"][v, t, 1]})
            d["heading"].append(rec["heading"][v, t]); d["existence"].append(rec["existence"][v, t]); d["timestep"].append(t)
        policy.update_state(vdd, v2e, t)
        out = policy.predict(vdd, gt_data_dict, sc["preproc"], None, v2e, t)
        assert out is vdd
        assert (policy.last_rtg_idx == g["rtg_idx"][t]).all(), t
        assert (policy.last_act_idx == g["act_idx"][t]).all(), t
        for v in v2e:
            if t >= cfg.nocturne.history_steps - 1:
                veh, (a, s_) = policy.act(Veh(v), t, vdd)
                assert veh.parked == (not rec["existence"][v, t]) and veh.steering == s_
                assert (veh.acceleration == a) if a > 0 else (veh.braked == abs(a))
                next_act[v] = (a, s_)
        port.apply_controls(ctx, t, v2e, next_act)
        for v in range(n):
            vdd[v]["acceleration"].append(rec["accel"][v, t]); vdd[v]["steering"].append(rec["steer"][v, t])
        sim.step(0.1)
    assert np.abs(rec["pos"][:, :T] - g["pos"][:, :T]).max() == 0.0  # the caller's simulator saw the reference's controls
    assert np.allclose(np.array([vdd[v]["rtgs"] for v in range(n)]), g["rtgs"][:, :T], atol=1e-12)


def test_evaluation_from_the_reference_file_layout(cfg, dev, tmp_path):
    """N4 on the GPU: scenes written in the reference's on-disk formats (test_filenames.pkl, Nocturne scenario JSONs,
    preprocess/test/*_physics.pkl; one scene without preprocessed data, which the reference skips without counting it)
    and the weights as a Lightning checkpoint; ``B200PolicyEvaluator(cfg, policy)`` reads them like
    ``PolicyEvaluator(cfg, policy)`` (policy_evaluator.py:33-41,436-464) and must give the metrics of the in-memory
    evaluation of the same scenes, and the partition JSON of policy_evaluator.py:578-593."""
    import json
    from ctrlsim_b200.checkpoint import save_checkpoint
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.model import DeviceModel
    from ctrlsim_b200.synth import make_scene, write_dataset
    from ctrlsim_b200.weights import make_weights
    scenes = [make_scene(60 + i, n_vehicles=5 + 2 * i, n_roads=1 + i % 2, n_chunks=3) for i in range(4)]
    paths = write_dataset(str(tmp_path), scenes)
    os.remove(os.path.join(paths["preprocess_dir"], "test", scenes[1]["name"] + "_physics.pkl"))
    ckpt = str(tmp_path / "run" / "model.ckpt")
    os.makedirs(os.path.dirname(ckpt))
    weights = make_weights(cfg, seed=8, still_bias=6.0)
    save_checkpoint(weights, ckpt, cfg)
    c = cfg.copy()
    c.eval = cfg.eval.copy()
    c.dataset_root, c.nocturne_waymo_val_folder = paths["dataset_root"], paths["nocturne_waymo_val_folder"]
    c.eval.num_files_to_evaluate, c.eval.multi_agent_eval_threshold = 3, 4
    steps = 12
    c.nocturne = cfg.nocturne.copy()
    c.nocturne.steps = steps
    pol = B200Policy(c, ckpt, DeviceModel.load_from_checkpoint(ckpt, c, dev), seed=2)
    ev = B200PolicyEvaluator(c, pol)
    m_files, lines = ev.evaluate_policy()
    assert ev.scene_ids == [0, 2, 3] and len(lines) == 9      # scene 1 has no *_physics.pkl: skipped, not counted
    kept = [scenes[0], scenes[2], scenes[3]]
    pol2 = B200Policy(c, "synthetic", DeviceModel(c, weights, dev), seed=2)
    ev2 = B200PolicyEvaluator(c, pol2, scenes=kept, scene_ids=[0, 2, 3])
    m_mem, _ = ev2.evaluate_policy()
    assert m_files == m_mem
    path = ev.write_partition_metrics()
    assert path == os.path.join(os.path.dirname(ckpt), "scene_results", "partition_0.json")
    saved = json.load(open(path))
    assert len(saved["collision"]) == 3 and len(saved["ade"]) == ev.batch.n_evaluated()
    assert abs(float(np.mean(saved["ade"])) - m_files["ade"]) < 1e-9 and abs(float(np.mean(saved["goal_success"])) - m_files["goal"]) < 1e-12


# ------------------------------------------------------------------ real-time rewards / decision transformer
def _dt_policy(cfg_dt, spec, dev):
    from ctrlsim_b200.evaluator import B200Policy
    p = cfg_dt.eval.policy
    return B200Policy(cfg_dt, "synthetic", _model(cfg_dt, spec, dev), use_rtg=p.use_rtg, predict_rtgs=p.predict_rtgs,
                      discretize_rtgs=p.discretize_rtgs, real_time_rewards=p.real_time_rewards,
                      max_return=p.max_return, min_return=p.min_return, name="dt", seed=0)


def test_dt_forward_matches_reference_logits(dev):
    """Decision-transformer network (cfgs/model/dt.yaml: continuous RTG inputs through Linear(1, H), (rtg, state, action)
    token order and its mask rule, no RTG head): the tokens the unmodified reference evaluator fed to its model at steps
    0 / 9 / 31 / 32 / 60 of the DT episode -> its recorded action logits."""
    from ctrlsim_b200.config import dt_config
    g, spec, _ = load_golden("dt")
    cfg_dt = dt_config()
    model = _model(cfg_dt, spec, dev)
    for t in spec["logit_steps"]:
        n_t = min(t + 1, 32)
        data = {k: g[f"in_{t}_{k}"][None] for k in ("agent_states", "agent_types", "goals", "actions", "road_points",
                                                    "road_types")}
        data["rtgs"] = g[f"in_{t}_rtgs_pass1"][None]
        data["timesteps"] = g[f"in_{t}_timesteps"][None][:, 0, :, 0]
        act_logits = model.forward_tokens_dt(data, n_t)
        n_real = int((g[f"in_{t}_agent_types"].sum(-1) > 0).sum())
        d_act = np.abs(act_logits[0, :n_real] - g[f"action_logits_{t}_0"][:n_real]).max()
        assert d_act < LOGIT_TOL, (t, d_act)


def test_dt_rollout_with_real_time_rewards_matches_reference(dev):
    """cfgs/policy/dt.yaml end to end: RTGs start at the maximum return and are decremented every step
    by the dense reward (signed distance to the road edges, nearest-vehicle distance, goal term; evaluator.py:106-140,
    policy_evaluator.py:123-149), one forward per focal group, action sampling only.  Free-running GPU episode vs the
    unmodified reference evaluator: sampled action bins bit for bit, states, the dense reward series and the tracked
    RTG series of all 90 steps, and the final metrics."""
    from ctrlsim_b200.config import dt_config
    from ctrlsim_b200.evaluator import B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    g, spec, ref_metrics = load_golden("dt")
    cfg_dt = dt_config()
    pol = _dt_policy(cfg_dt, spec, dev)
    ev = B200PolicyEvaluator(cfg_dt, pol, scenes=[make_scene(**spec["scene"])])
    b = ev.build_batch(eval_threshold=64)
    ev.rollout(b)
    torch.cuda.synchronize()
    tr = b.trace()
    n = g["pos"].shape[0]
    act = tr["tr_act_idx"][0, :n, :90].T.astype(np.int64)
    bad = np.argwhere(act != g["act_idx"])
    T = int(bad[:, 0].min()) if len(bad) else 90
    assert T >= 60, ("an action draw differs early", T, bad[:4].tolist())
    assert (tr["tr_rtg_idx"][0, :n] == -1).all()  # nothing is sampled for the RTGs
    ex = g["existence"][:, :T + 1].astype(bool)
    assert (tr["tr_exist"][0, :n, :T + 1] == g["existence"][:, :T + 1]).all()
    dpos = np.abs(tr["tr_pos"][0, :n, :T + 1].astype(np.float64) - g["pos"][:, :T + 1])[ex].max()
    dhead = np.abs(tr["tr_heading"][0, :n, :T + 1].astype(np.float64) - g["heading"][:, :T + 1])[ex].max()
    assert dpos == 0.0 and dhead == 0.0, (dpos, dhead)
    # dense reward of the states 0..T and the RTG series it drives (evaluated vehicles; others keep their start value)
    dd = np.abs(tr["tr_dense"][0, :n, :T + 1] - g["dense_reward"][:, :T + 1])[ex].max()
    assert dd < 1e-9, dd
    Tr = min(T + 1, 90)
    dr = np.abs(tr["rt_rtg"][0, :n, :Tr] - g["rtgs"][:, :Tr])[g["existence"][:, :Tr].astype(bool)].max()
    assert dr < 1e-9, dr
    if T == 90:
        m = ev.metrics_from_summary(ev.summarize(b))
        for k, v in ref_metrics.items():
            assert abs(m[k] - v) < 1e-4 * max(1.0, abs(v)), (k, m[k], v)


@pytest.mark.parametrize("mode", ["data", "min_return"])
def test_ctrl_sim_network_with_tracked_rtgs_matches_oracle_port(cfg, dev, mode):
    """real_time_rewards with the CtRL-Sim network (predict_rtgs=False, discretize_rtgs=True; policies/policy.py:89-118):
    the tracked RTG series - started from the logged returns of the *_physics.pkl ('data') or from min_return - is
    clip-normalised and discretised into the RTG embedding bins, one forward per group, actions sampled.  GPU == the
    oracle port, scene by scene: action bins, positions, dense reward, RTG series."""
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    from ctrlsim_b200.model import DeviceModel
    from oracle.model_port import ModelPort
    from oracle.policy_port import RolloutPort
    weights = make_weights(cfg, seed=5, still_bias=2.0)
    scenes = [make_scene(70 + i, n_vehicles=5 + 3 * i, n_roads=2, n_chunks=3) for i in range(2)]
    rng = np.random.default_rng(12)
    for sc in scenes:  # a logged episode with non-trivial returns, so that 'data' starts from different RTGs per vehicle
        pre = sc["preproc"]
        n_obj = pre["num_agents"]
        pre["ag_data"][:, :, -1] = rng.random((n_obj, 90)) < 0.9
        pre["ag_rewards"][:, :, 0] = rng.random((n_obj, 90)) < 0.05
        pre["ag_rewards"][:, :, 6] = rng.random((n_obj, 90)) < 0.02
        pre["ag_rewards"][:, :, 7] = rng.random((n_obj, 90)) < 0.02
        pre["veh_edge_dist_rewards"][:] = rng.normal(size=(n_obj, 90)) * 0.2
        pre["veh_veh_dist_rewards"][:] = rng.random((n_obj, 90))
    steps = 5
    kw = dict(predict_rtgs=False, discretize_rtgs=True, real_time_rewards=True, min_return=(mode == "min_return"))
    pol = B200Policy(cfg, "synthetic", DeviceModel(cfg, weights, dev), seed=4, **kw)
    ev = B200PolicyEvaluator(cfg, pol, scenes=scenes)
    b = ev.build_batch(eval_threshold=64)
    ev.rollout(b, max_steps=steps)
    tr = b.trace()
    port = RolloutPort(cfg, ModelPort(cfg, weights), seed=4, eval_threshold=64, **kw)
    for s, sc in enumerate(scenes):
        rec = port.run_scene(s, sc["json"], sc["preproc"], max_steps=steps)
        n = rec["n"]
        assert (tr["tr_act_idx"][s, :n, :steps].T == rec["act_idx"][:steps]).all()
        assert np.abs(tr["tr_pos"][s, :n, :steps] - rec["pos"][:, :steps]).max() < POS_TOL
        assert np.abs(tr["tr_dense"][s, :n, :steps] - rec["dense_reward"][:, :steps]).max() < 1e-9
        assert np.abs(tr["rt_rtg"][s, :n, :steps] - rec["rtgs"][:, :steps]).max() < 1e-9


def test_reference_signature_adapter_with_real_time_rewards(dev):
    """SURVEY 8(b) x 8(f) N1: the per-scene adapter in cfgs/policy/dt.yaml mode.  The stock evaluator tracks the RTGs
    itself and hands them over in vehicle_data_dict['rtgs'][t] (policy_evaluator.py:123-149, policies/policy.py:91-93);
    here the reference's recorded world of the DT fixture (states, applied controls, tracked RTGs) is replayed through
    reset / update_state / predict / act, across the switch to the sliding window: every sampled action bin must be the
    one the unmodified reference drew, and predict() must leave the RTG series alone."""
    from ctrlsim_b200.config import dt_config
    from ctrlsim_b200.policy_adapter import B200AutoregressivePolicy
    from ctrlsim_b200.synth import make_scene
    from oracle.policy_port import RolloutPort
    g, spec, _ = load_golden("dt")
    cfg_dt = dt_config()
    sc = make_scene(**spec["scene"])
    T = 36
    kd = {"next_acceleration": "next_acceleration", "next_steering": "next_steering", "rtgs": "rtgs"}
    td = {"tilt": True, "goal_tilt": 0, "veh_veh_tilt": 0, "veh_edge_tilt": 0}
    p = cfg_dt.eval.policy
    policy = B200AutoregressivePolicy(cfg_dt, "synthetic", _model(cfg_dt, spec, dev), p.use_rtg, p.predict_rtgs, p.discretize_rtgs,
                                      p.real_time_rewards, p.privileged_return, p.max_return, p.min_return, kd, td, "dt", 1.0,
                                      False, 0.8, seed=0)
    ctx = RolloutPort(cfg_dt, None).setup_scene(0, sc["json"])
    n, gt = ctx["n"], ctx["gt"]
    v2e = [int(v) for v in g["evaluated"]]
    gt_data_dict = {v: {"traj": [list(gt[v, t]) for t in range(91)]} for v in range(n)}
    vdd = {v: {"position": [], "velocity": [], "heading": [], "existence": [], "acceleration": [], "steering": [], "timestep": [],
               "rtgs": [], "goal_position": {"x": ctx["goal"][v, 0], "y": ctx["goal"][v, 1]}, "goal_heading": ctx["goal"][v, 2],
               "goal_speed": ctx["goal"][v, 3], "length": float(ctx["parsed"]["size"][v, 0]), "width": float(ctx["parsed"]["size"][v, 1]),
               "type": "vehicle", "next_acceleration": 0.0, "next_steering": 0.0} for v in range(n)}
    policy.scene_index = -1
    policy.reset(vdd)
    undisc = lambda a: ((a // 50) * 2.0 * 10 / 19 - 10, (a % 50) * 2.0 * 0.7 / 49 - 0.7)  # dataset.py undiscretize_actions
    for t in range(T):
        for v in range(n):
            d = vdd[v]
            d["position"].append({"x": g["pos"][v, t, 0], "y": g["pos"][v, t, 1]})
            d["velocity"].append({"x": g["vel"][v, t, 0], "y": g["vel"][v, t, 1]})
            d["heading"].append(g["heading"][v, t]); d["existence"].append(g["existence"][v, t]); d["timestep"].append(t)
            d["rtgs"].append(g["rtgs"][v, t].copy())
        policy.update_state(vdd, v2e, t)
        out = policy.predict(vdd, gt_data_dict, sc["preproc"], None, v2e, t)
        assert out is vdd and all(len(vdd[v]["rtgs"]) == t + 1 for v in range(n))
        assert (policy.last_act_idx == g["act_idx"][t]).all(), (t, policy.last_act_idx, g["act_idx"][t])
        for v in v2e:
            if g["act_idx"][t, v] >= 0:
                a, s_ = undisc(int(g["act_idx"][t, v]))
                assert abs(vdd[v]["next_acceleration"] - a) < 1e-5 and abs(vdd[v]["next_steering"] - s_) < 1e-6
        for v in range(n):
            vdd[v]["acceleration"].append(g["accel"][v, t]); vdd[v]["steering"].append(g["steer"][v, t])


def test_dt_multi_scene_rollout_matches_oracle_port(dev):
    """Decision-transformer mode beyond the single fixture scene: a crowded scene (30 vehicles > the 24-agent cap, so
    relevant-agent selection and several focal groups per step are exercised with the (rtg, state, action) order) and a
    small one in ONE batch, min_return start (evaluated vehicles (0, -10, -10), the others (10, 90, 90)), against the
    oracle port scene by scene: action bins, positions, dense reward, tracked RTGs."""
    from ctrlsim_b200.config import dt_config
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    from ctrlsim_b200.model import DeviceModel
    from oracle.model_port import ModelPort
    from oracle.policy_port import RolloutPort
    cfg_dt = dt_config()
    weights = make_weights(cfg_dt, seed=7, still_bias=2.0)
    scenes = [make_scene(80, n_vehicles=30, n_roads=3, n_chunks=3), make_scene(81, n_vehicles=8, n_roads=2, n_chunks=3)]
    steps = 3
    kw = dict(predict_rtgs=False, discretize_rtgs=False, real_time_rewards=True, min_return=True)
    pol = B200Policy(cfg_dt, "synthetic", DeviceModel(cfg_dt, weights, dev), seed=2, name="dt", **kw)
    ev = B200PolicyEvaluator(cfg_dt, pol, scenes=scenes)
    b = ev.build_batch(eval_threshold=8)
    ev.rollout(b, max_steps=steps)
    tr = b.trace()
    port = RolloutPort(cfg_dt, ModelPort(cfg_dt, weights), seed=2, eval_threshold=8, **kw)
    for s, sc in enumerate(scenes):
        rec = port.run_scene(s, sc["json"], sc["preproc"], max_steps=steps)
        n = rec["n"]
        assert sorted(rec["evaluated"]) == b.evaluated_ids[s]
        assert (tr["tr_act_idx"][s, :n, :steps].T == rec["act_idx"][:steps]).all()
        assert np.abs(tr["tr_pos"][s, :n, :steps] - rec["pos"][:, :steps]).max() < POS_TOL
        assert np.abs(tr["tr_dense"][s, :n, :steps] - rec["dense_reward"][:, :steps]).max() < 1e-9
        assert np.abs(tr["rt_rtg"][s, :n, :steps] - rec["rtgs"][:, :steps]).max() < 1e-9
    assert {tuple(v) for v in tr["rt_rtg"][0, :30, 0].tolist()} == {(0.0, -10.0, -10.0), (10.0, 90.0, 90.0)}


def test_dt_as_shipped_rollout_matches_reference(dev):
    """The decision-transformer baseline with the switches Hydra composes from the reference's files AS SHIPPED:
    cfgs/policy/dt.yaml:11 spells `use_rtgs`, so use_rtg stays False and the network is fed RTG (0, 0, 0) while the
    evaluator keeps computing the dense reward.  34 free-running steps (two of them with the sliding window) vs the
    unmodified reference evaluator: sampled action bins, states bit for bit, dense reward; the policy's RTG buffer is zero."""
    from ctrlsim_b200.config import dt_config
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    g, spec, _ = load_golden("dt_as_shipped")
    cfg_dt = dt_config(as_shipped=True)
    cfg_dt.nocturne = cfg_dt.nocturne.copy()
    cfg_dt.nocturne.steps = spec["steps"]
    p = cfg_dt.eval.policy
    assert p.use_rtg is False
    pol = B200Policy(cfg_dt, "synthetic", _model(cfg_dt, spec, dev), use_rtg=p.use_rtg, predict_rtgs=p.predict_rtgs,
                     discretize_rtgs=p.discretize_rtgs, real_time_rewards=p.real_time_rewards, max_return=p.max_return,
                     min_return=p.min_return, name="dt", seed=0)
    ev = B200PolicyEvaluator(cfg_dt, pol, scenes=[make_scene(**spec["scene"])])
    b = ev.build_batch(eval_threshold=64)
    ev.rollout(b)
    tr = b.trace()
    T, n = spec["steps"], g["pos"].shape[0]
    assert (tr["tr_act_idx"][0, :n, :T].T == g["act_idx"][:T]).all()
    ex = g["existence"][:, :T + 1].astype(bool)
    assert (tr["tr_pos"][0, :n, :T + 1].astype(np.float64)[ex] == g["pos"][:, :T + 1][ex]).all()
    assert (tr["tr_heading"][0, :n, :T + 1].astype(np.float64)[ex] == g["heading"][:, :T + 1][ex]).all()
    assert np.abs(tr["tr_dense"][0, :n, :T + 1] - g["dense_reward"][:, :T + 1])[ex].max() < 1e-9
    assert (tr["rt_rtg"][0] == 0).all()
    d_act = np.abs(pol.model.forward_tokens_dt({**{k: g[f"in_33_{k}"][None] for k in ("agent_states", "agent_types", "goals", "actions", "road_points", "road_types")},
                                                "rtgs": g["in_33_rtgs_pass1"][None], "timesteps": g["in_33_timesteps"][None][:, 0, :, 0]}, 32)[0, :n]
                   - g["action_logits_33_0"][:n]).max()
    assert d_act < LOGIT_TOL, d_act


def test_pipelined_evaluation_equals_one_batch(cfg, dev):
    """evaluate_policy(sub_batch_scenes=2) - next sub-batch parsed and uploaded on a side stream while the current one
    rolls out - gives the metrics and, scene by scene, the traces of the one-batch evaluation (whole 90-step episodes)."""
    from ctrlsim_b200.evaluator import B200Policy, B200PolicyEvaluator
    from ctrlsim_b200.synth import make_scene
    from ctrlsim_b200.weights import make_weights
    from ctrlsim_b200.model import DeviceModel
    scenes = [make_scene(500 + i, n_vehicles=6 + 3 * i, n_roads=2, n_chunks=3) for i in range(5)]
    ids = [500 + i for i in range(5)]
    model = DeviceModel(cfg, make_weights(cfg, seed=4, still_bias=3.0), dev)
    one = B200PolicyEvaluator(cfg, B200Policy(cfg, "synthetic", model, seed=1), scenes=scenes, scene_ids=ids)
    m1, _ = one.evaluate_policy(keep_traces=True)
    pip = B200PolicyEvaluator(cfg, B200Policy(cfg, "synthetic", model, seed=1), scenes=scenes, scene_ids=ids)
    m2, _ = pip.evaluate_policy(sub_batch_scenes=2, keep_traces=True)
    assert len(pip.traces) == 3 and pip.n_evaluated == one.batch.n_evaluated()
    for k in m1:
        assert abs(m1[k] - m2[k]) <= 1e-12 * max(1.0, abs(m1[k])), (k, m1[k], m2[k])
    assert np.array_equal(one.last_summary[8:], pip.last_summary[8:])  # histograms: integers, exact
    s = 0
    for tr in pip.traces:
        for j in range(tr["tr_pos"].shape[0]):
            n = int(tr["n_veh"][j])
            for f in ("tr_act_idx", "tr_rtg_idx", "tr_pos", "tr_heading", "tr_exist"):
                assert np.array_equal(tr[f][j, :n], one.traces[0][f][s, :n]), (s, f)
            s += 1
    assert s == 5
