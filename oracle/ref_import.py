"""TEST INFRASTRUCTURE — import the LIVE reference ``TrajectoryHead`` in the build container.

The reference (/root/reference, read-only, Python only) cannot be imported as-is:
diffusers, timm, nuplan, pytorch_lightning, hydra, omegaconf, shapely,
pyquaternion, ray and matplotlib are absent and there is no network.  This module
installs a ``sys.meta_path`` finder that answers those top-level packages with
``MagicMock`` modules, pre-seeds the three pieces the head really needs
(``TrajectorySampling``, ``LightningModule``, ``DDIMScheduler`` — the latter with
oracle/ddim.py, "parity unpinned") and imports
``navsim.agents.diffusiondrive.transfuser_model_v2`` from where it lies.

It is used ONLY by oracle/make_golden.py and by the optional container-only tests
to pin oracle/head_oracle.py against the real module.  /root/reference does not
exist on the GPU box; nothing that runs there imports this file.
"""
from __future__ import annotations

import dataclasses
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("DDH_REFERENCE_ROOT", "/root/reference")

_STUB_TOP = ("nuplan", "shapely", "timm", "pytorch_lightning", "pyquaternion", "diffusers",
             "matplotlib", "hydra", "omegaconf", "ray", "cv2", "PIL", "torchvision")


class _Loader(importlib.abc.Loader):
    def create_module(self, spec):
        m = MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        pass


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in getattr(self, "names", _STUB_TOP):
            return importlib.machinery.ModuleSpec(name, _Loader(), is_package=True)
        return None


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "navsim", "agents", "diffusiondrive"))


_cached = None


def load_reference():
    """Returns (TrajectoryHead, TransfuserConfig, module) of the live reference."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    import torch.nn as nn

    from .ddim import DDIMSchedulerRestated

    # stub only what is really absent from this image
    import importlib.util
    missing = tuple(n for n in _STUB_TOP if importlib.util.find_spec(n) is None)
    finder = _Finder()
    finder.names = missing
    sys.meta_path.insert(0, finder)

    @dataclasses.dataclass(frozen=True)
    class TrajectorySampling:
        num_poses: int = None
        time_horizon: float = None
        interval_length: float = None

        def __post_init__(self):
            if self.num_poses is None:
                object.__setattr__(self, "num_poses",
                                   int(self.time_horizon / self.interval_length))

    ts_mod = types.ModuleType("nuplan.planning.simulation.trajectory.trajectory_sampling")
    ts_mod.TrajectorySampling = TrajectorySampling
    sys.modules[ts_mod.__name__] = ts_mod

    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = nn.Module
    pl.Callback = object
    pl.__path__ = []
    sys.modules["pytorch_lightning"] = pl

    dsch = types.ModuleType("diffusers.schedulers")
    dsch.DDIMScheduler = DDIMSchedulerRestated
    dmod = types.ModuleType("diffusers")
    dmod.schedulers = dsch
    dmod.__path__ = []
    sys.modules["diffusers"] = dmod
    sys.modules["diffusers.schedulers"] = dsch

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mod = importlib.import_module("navsim.agents.diffusiondrive.transfuser_model_v2")
    cfg_mod = importlib.import_module("navsim.agents.diffusiondrive.transfuser_config")
    _cached = (mod.TrajectoryHead, cfg_mod.TransfuserConfig, mod)
    return _cached


def build_reference_head(state_dict, anchors_npy_path: str, num_layers: int = 2):
    """Instantiate the live reference head and load ``state_dict`` into it.

    ``num_layers != 2`` (stress config) swaps ``diff_decoder`` for a deeper stack of the
    reference's own layer class, because the depth is a literal (:476)."""
    TrajectoryHead, TransfuserConfig, mod = load_reference()
    cfg = TransfuserConfig()
    cfg.plan_anchor_path = anchors_npy_path
    head = TrajectoryHead(num_poses=8, d_ffn=cfg.tf_d_ffn, d_model=cfg.tf_d_model,
                          plan_anchor_path=anchors_npy_path, config=cfg)
    if num_layers != 2:
        layer = mod.CustomTransformerDecoderLayer(num_poses=8, d_model=cfg.tf_d_model,
                                                  d_ffn=cfg.tf_d_ffn, config=cfg)
        head.diff_decoder = mod.CustomTransformerDecoder(layer, num_layers)
    missing, unexpected = head.load_state_dict(state_dict, strict=True)
    assert not missing and not unexpected
    return head.eval(), cfg
