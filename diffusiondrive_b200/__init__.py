"""B200-native DiffusionDrive planning head (TrajectoryHead inference path).

Public surface:
  TrajectoryHead   drop-in for navsim.agents.diffusiondrive.transfuser_model_v2.TrajectoryHead
  HeadConfig       the slice of TransfuserConfig the head reads
  synth            synthetic weights / anchors / features / noise (seeded)
  parallel         scene sharding across the GPUs of one box
"""
from .config import HeadConfig, roll_timesteps  # noqa: F401
from .trajectory_head import TrajectoryHead, ddim_alphas_cumprod  # noqa: F401
from . import synth  # noqa: F401

__all__ = ["TrajectoryHead", "HeadConfig", "roll_timesteps", "ddim_alphas_cumprod", "synth"]
