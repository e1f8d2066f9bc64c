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
    missing = tuple(n for n in _STUB_TOP if n not in sys.modules and importlib.util.find_spec(n) is None)
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


# ------------------------------------------------------------------------------------------
# Full model (BASELINE configs[3]): V2TransfuserModel needs ``timm.create_model(...,
# features_only=True)``; timm is absent, so a functional stand-in built on torchvision's resnet34
# is installed: an ``nn.ModuleDict`` with timm's child names (conv1, bn1, act1, maxpool,
# layer1..4), ``return_layers`` (5 entries, so the reference skips the stem: start_index = 1,
# transfuser_backbone.py:62-65) and ``feature_info.info``.
def _install_timm_shim():
    import torch.nn as nn
    import torchvision

    class FeatureNet(nn.ModuleDict):
        def __init__(self, in_chans):
            m = torchvision.models.resnet34(weights=None)
            if in_chans != 3:
                m.conv1 = nn.Conv2d(in_chans, 64, 7, 2, 3, bias=False)
            super().__init__({"conv1": m.conv1, "bn1": m.bn1, "act1": m.relu, "maxpool": m.maxpool,
                              "layer1": m.layer1, "layer2": m.layer2, "layer3": m.layer3,
                              "layer4": m.layer4})
            self.return_layers = {"act1": "0", "layer1": "1", "layer2": "2", "layer3": "3", "layer4": "4"}
            self.feature_info = types.SimpleNamespace(info=[
                {"num_chs": c, "reduction": r} for c, r in zip((64, 64, 128, 256, 512), (2, 4, 8, 16, 32))])

    def create_model(name, pretrained=False, features_only=True, in_chans=3, **kw):
        assert name == "resnet34" and features_only
        return FeatureNet(in_chans)

    timm = types.ModuleType("timm")
    timm.create_model = create_model
    timm.__spec__ = importlib.machinery.ModuleSpec("timm", None)
    sys.modules["timm"] = timm


def build_reference_model(state_dict, anchors_npy_path: str):
    """The live reference ``V2TransfuserModel`` (transfuser_model_v2.py:19-162) with ``state_dict``."""
    _install_timm_shim()
    for name in [n for n in sys.modules if n.startswith("navsim.agents.diffusiondrive.transfuser_backbone")]:
        del sys.modules[name]
    global _cached
    _cached = None
    _, TransfuserConfig, mod = load_reference()
    import importlib
    bb = importlib.import_module("navsim.agents.diffusiondrive.transfuser_backbone")
    import timm as shim
    bb.timm = shim
    mod.TransfuserBackbone = bb.TransfuserBackbone
    cfg = TransfuserConfig()
    cfg.plan_anchor_path = anchors_npy_path
    model = mod.V2TransfuserModel(cfg)
    missing, unexpected = model.load_state_dict(state_dict, strict=True)
    assert not missing and not unexpected
    return model.eval(), cfg
