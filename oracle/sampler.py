"""ORACLE (test infrastructure only): CPU statement of the explicit categorical sampler ("sampler contract v1").

The reference draws every RTG component and every action with ``torch.multinomial`` on a softmax output, one vehicle
at a time (policies/policy.py:122-127, policies/autoregressive_policy.py:233-236).  Which index comes out depends on
the position of torch's global RNG stream, i.e. on call order - no batched GPU kernel can reproduce that stream.
Both sides therefore use the following explicit, order-independent sampler instead (installed into the reference by
oracle/ref_harness.py); it is a faithful multinomial draw from softmax(x) up to a 2^-30 quantisation of the weights:

  x_i   RTG:    fp32( fp64(logit_i) + tilt * lin_i ),  lin = np.linspace(0, 1, 350)   (dataset.py:342-348; the
                reference adds the float64 tilt to the float32 logits, policy.py:117-126)
        action: fp32( logit_i / temperature )                                         (autoregressive_policy.py:233)
  d_i = max(x_i - max_j x_j, -80)                                  (fp32)
  e_i = exp_spec(d_i)    a fixed fp32 algorithm, every multiply/add rounded separately (no FMA), bit-reproducible
  w_i = floor(e_i * 2^30)                                          (integer; the arg-max bin has w = 2^30)
  r   = Philox4x32-10(key=seed, counter=(scene, agent, step, component))  -> 64 random bits
  idx = min{ i : w_0 + ... + w_i > mulhi64(r, sum_j w_j) }         (integer prefix sums: any scan order agrees)

Nucleus (top-p) sampling of the action (autoregressive_policy.py:216-230, off in every reference config) is stated on the
same integer weights: order the categories by (w descending, index ascending); with cum_k the inclusive prefix sums
in that order keep category k iff k == 0 or float64(cum_{k-1}) < float64(p) * float64(sum_j w_j) - the reference's
"cum_probs < p shifted right by one" - and draw idx = the first KEPT i, in index order, whose inclusive prefix over
the kept weights exceeds mulhi64(r, sum_kept w).  (The reference takes the prefix on fp32 probabilities; the two sets
differ only when a cumulative probability lies within 2^-30 of p.)

component: 0/1/2 = RTG goal / veh / road, 3 = action.  `scene` is the global scene index, `agent` the vehicle's
index in the scenario, so results do not depend on how scenes are sharded over GPUs.
"""
from __future__ import annotations

import numpy as np

from ctrlsim_b200.philox import philox4x32

_LOG2E = np.float32(1.4426950408889634)
_LN2_HI = np.float32(0.693145751953125)
_LN2_LO = np.float32(1.428606765330187e-06)
_C = [np.float32(c) for c in (1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0)]

COMP_RTG_GOAL, COMP_RTG_VEH, COMP_RTG_ROAD, COMP_ACTION = 0, 1, 2, 3


def exp_spec(d):
    """fp32 exp for d in [-80, 0]: Cody-Waite reduction + degree-6 Horner, each op rounded to fp32."""
    d = np.asarray(d, dtype=np.float32)
    k = np.rint(d * _LOG2E).astype(np.float32)
    r = (d - k * _LN2_HI).astype(np.float32)
    r = (r - k * _LN2_LO).astype(np.float32)
    p = np.full_like(r, _C[0])
    for c in _C[1:]:
        p = (p * r).astype(np.float32)
        p = (p + c).astype(np.float32)
    return np.ldexp(p, k.astype(np.int32)).astype(np.float32)


def weights_from_x(x):
    x = np.asarray(x, dtype=np.float32)
    d = np.maximum((x - x.max()).astype(np.float32), np.float32(-80.0))
    e = exp_spec(d)
    return np.floor(e.astype(np.float64) * float(1 << 30)).astype(np.uint64)


def random_bits(seed: int, scene: int, agent: int, step: int, comp: int) -> int:
    ctr = np.array([scene, agent, step, comp], dtype=np.uint32)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    o = philox4x32(ctr, key)
    return int(o[0]) | (int(o[1]) << 32)


def sample_from_x(x, seed, scene, agent, step, comp) -> int:
    w = weights_from_x(x)
    total = int(w.sum())
    target = (random_bits(seed, scene, agent, step, comp) * total) >> 64
    return int(np.searchsorted(np.cumsum(w), np.uint64(target), side="right"))


def nucleus_keep(w, p: float):
    """Boolean keep mask of the top-p set on integer weights w (definition above; straightforward sort)."""
    w = np.asarray(w, dtype=np.uint64)
    n = len(w)
    order = np.lexsort((np.arange(n), -w.astype(np.int64)))  # w descending, index ascending
    cum = np.cumsum(w[order])
    thr = float(p) * float(int(w.sum()))
    keep_sorted = np.ones(n, bool)
    keep_sorted[1:] = cum[:-1].astype(np.float64) < thr
    keep = np.zeros(n, bool)
    keep[order] = keep_sorted
    return keep


def sample_from_x_nucleus(x, p, seed, scene, agent, step, comp) -> int:
    w = weights_from_x(x)
    w = np.where(nucleus_keep(w, p), w, np.uint64(0))
    total = int(w.sum())
    target = (random_bits(seed, scene, agent, step, comp) * total) >> 64
    return int(np.searchsorted(np.cumsum(w), np.uint64(target), side="right"))


def rtg_x(logits_f32, tilt: float, n_bins: int = 350):
    lin = np.linspace(0.0, 1.0, n_bins)
    return (np.asarray(logits_f32, dtype=np.float32).astype(np.float64) + tilt * lin).astype(np.float32)


def action_x(logits_f32, temperature: float):
    return (np.asarray(logits_f32, dtype=np.float32) / np.float32(temperature)).astype(np.float32)
