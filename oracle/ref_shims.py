"""ORACLE (test infrastructure only): import the UNMODIFIED reference python packages from /root/reference.

Only usable where /root/reference exists (the build container).  It is used to (1) validate oracle/port_*.py and
(2) generate the committed fixtures under tests/golden/ (oracle/make_golden.py).  Nothing here travels to the GPU box
except the fixtures it produced.

The reference imports hydra, omegaconf, pytorch_lightning, torch_geometric, torch_scatter, matplotlib, imageio,
pyvirtualdisplay ... none of which are installed and none of which matter on the evaluation path.  A meta-path
finder fabricates stub modules for those roots; five behaviours are filled in for real:
hydra.main, torch_geometric.data.{Dataset,HeteroData}, torch_geometric.nn.conv.MessagePassing,
pytorch_lightning.LightningModule.  ``modules.diffusion_guidance`` is absent from the reference itself
(modules/diffusion.py:14) and is stubbed too.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

import torch
import torch.nn as nn

REF = os.environ.get("CTRLSIM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
STUB_ROOTS = {"hydra", "omegaconf", "pyvirtualdisplay", "matplotlib", "imageio", "moviepy", "torch_geometric",
              "torch_scatter", "pytorch_lightning", "torch_ema", "seaborn", "cv2", "wandb"}


class _Any:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return a[0] if (len(a) == 1 and callable(a[0]) and not k) else _Any()

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Any()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _StubModule(types.ModuleType):
    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        v = type(n, (_Any,), {})
        setattr(self, n, v)
        return v


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, m):
        pass


_installed = False


def install():
    """Idempotent. After this, `import policies, evaluators, models, nocturne` resolve to the reference."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference tree not found at {REF}")
    sys.meta_path.insert(0, _Finder())

    import hydra
    hydra.main = lambda *a, **k: (lambda f: f)
    import torch_geometric.data as tgd
    import torch_geometric.nn.conv as tgc

    class Dataset:
        def __init__(self, *a, **k):
            pass

        def __getitem__(self, i):
            return self.get(i)

        def __len__(self):
            return self.len()

    class _Store(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__

    class HeteroData:
        def __init__(self, d=None):
            object.__setattr__(self, "_s", {})
            for k, v in (d or {}).items():
                self[k] = v

        def __getitem__(self, k):
            return self._s[k]

        def __setitem__(self, k, v):
            self._s[k] = _Store(v) if isinstance(v, dict) else v

        def cuda(self, *a, **k):
            return self

        def to(self, *a, **k):
            return self

    tgd.Dataset = Dataset
    tgd.HeteroData = HeteroData
    tgc.MessagePassing = nn.Module
    import pytorch_lightning as pl

    class LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    class LightningDataModule:
        def __init__(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    pl.LightningDataModule = LightningDataModule
    dg = types.ModuleType("modules.diffusion_guidance")
    dg.n_step_guided_p_sample = dg.GoalGuide = dg.CollisionGuide = _Any
    sys.modules["modules.diffusion_guidance"] = dg
    if not torch.cuda.is_available():  # the reference hard-codes .cuda() (policy.py:118,120; autoregressive_policy.py:183)
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.cuda.manual_seed = lambda *a, **k: None
    sys.path.insert(0, os.path.join(HERE, "_ref"))  # nocturne_cpp*.so built by oracle/Makefile
    sys.path.insert(0, REF)
    # reference `datasets/` has no __init__.py and would lose to the HuggingFace package of the same name
    ds = types.ModuleType("datasets")
    ds.__path__ = [os.path.join(REF, "datasets")]
    sys.modules["datasets"] = ds
    _installed = True
