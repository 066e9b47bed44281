"""Weight import from the reference's on-disk checkpoint format.

The reference restores its model with ``CtRLSim.load_from_checkpoint(model_path)`` (eval_sim.py:52): a PyTorch
Lightning ``.ckpt`` is a ``torch.save`` dict whose ``state_dict`` holds the ``encoder.*`` / ``decoder.*`` tensors of
models/ctrl_sim.py:29-30 (and, when training used the EMA callback, a parallel set the reference ignores at load).
``load_checkpoint`` reads such a file, checks every tensor the rollout needs against ``weights.param_spec`` (name and
shape, nothing silently skipped) and returns the float32 numpy dict ``model.DeviceModel`` consumes.
``save_checkpoint`` writes the same layout (used by the tests and to hand weights to the reference).
"""
from __future__ import annotations

import pickle

from collections import OrderedDict

import numpy as np
import torch

from .weights import param_spec

# tensors of the reference module that the inference path never reads (modules/decoder.py:70-72: the future-state head
# is evaluated but its output is dropped)
_UNUSED_PREFIXES = ("decoder.predict_future_states",)
_WRAPPER_PREFIXES = ("model.", "module.", "_orig_mod.")


class CheckpointError(ValueError):
    pass


def _strip(key: str) -> str:
    changed = True
    while changed:
        changed = False
        for p in _WRAPPER_PREFIXES:
            if key.startswith(p):
                key, changed = key[len(p):], True
    return key


def check_state_dict(state_dict, cfg, strict_unexpected: bool = False):
    """Validate names and shapes against the architecture ``cfg`` describes. Returns an OrderedDict in spec order
    (float32 numpy). Raises CheckpointError naming every missing or mis-shaped tensor."""
    have = {_strip(k): v for k, v in state_dict.items()}
    out, problems = OrderedDict(), []
    wanted = set()
    for key, shape, _kind in param_spec(cfg):
        wanted.add(key)
        if key not in have:
            if not key.startswith(_UNUSED_PREFIXES):
                problems.append(f"missing '{key}' {tuple(shape)}")
            continue
        v = have[key]
        a = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        if tuple(a.shape) != tuple(shape):
            problems.append(f"'{key}' has shape {tuple(a.shape)}, the configuration implies {tuple(shape)}")
            continue
        if not np.issubdtype(a.dtype, np.floating):
            problems.append(f"'{key}' has dtype {a.dtype}")
            continue
        out[key] = np.ascontiguousarray(a, dtype=np.float32)
    extra = sorted(k for k in have if k not in wanted and "causal_mask" not in k)
    if extra and strict_unexpected:
        problems.append("unexpected tensors: " + ", ".join(extra[:8]) + (" ..." if len(extra) > 8 else ""))
    if problems:
        raise CheckpointError("checkpoint does not match the model configuration:\n  " + "\n  ".join(problems))
    return out


def load_checkpoint(path: str, cfg, strict_unexpected: bool = False, trusted: bool = False):
    """Read a Lightning ``.ckpt`` (or a bare ``state_dict`` file) -> OrderedDict[str, np.float32 array].

    The file is read with ``weights_only=True``.  Lightning checkpoints pickle their hyper-parameters (an omegaconf tree
    in the reference, train.py:46), which that mode refuses; ``trusted=True`` - the caller vouches for the file, as the
    reference's ``load_from_checkpoint`` implicitly does - falls back to full unpickling for exactly that error.  Any
    other failure (missing file, corrupt archive) propagates unchanged."""
    try:
        blob = torch.load(path, map_location="cpu", weights_only=True)
    except pickle.UnpicklingError as e:
        if not trusted:
            raise CheckpointError(f"{path}: holds pickled objects besides tensors ({e}); pass trusted=True to unpickle a "
                                  "checkpoint you trust") from e
        blob = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(blob, dict) and "state_dict" in blob:
        blob = blob["state_dict"]
    if not isinstance(blob, dict) or not blob:
        raise CheckpointError(f"{path}: no state_dict found")
    return check_state_dict(blob, cfg, strict_unexpected)


def save_checkpoint(state_dict, path: str, cfg=None, lightning_version: str = "2.0.0"):
    """Write ``state_dict`` in the Lightning layout ``load_from_checkpoint`` reads."""
    sd = OrderedDict((k, torch.from_numpy(np.array(v, dtype=np.float32, copy=True))) for k, v in state_dict.items())
    blob = {"state_dict": sd, "epoch": 0, "global_step": 0, "pytorch-lightning_version": lightning_version}
    if cfg is not None:
        blob["hyper_parameters"] = {"cfg": _plain(cfg)}
    torch.save(blob, path)


def _plain(x):
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    return x
