"""Static description of the planning-head path.

The reference spreads these numbers over ``TransfuserConfig``
(navsim/agents/diffusiondrive/transfuser_config.py:11-149) and literals inside
``TrajectoryHead`` (navsim/agents/diffusiondrive/transfuser_model_v2.py:431-641).
Everything the hot path reads is gathered here so it can cross the C ABI as one
plain struct (``ddh_shape`` in include/ddh.h).
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass(frozen=True)
class HeadConfig:
    """Duck-type stand-in for the slice of ``TransfuserConfig`` the head reads.

    Attribute names follow transfuser_config.py (tf_d_model :74, tf_d_ffn :75,
    tf_num_head :77, tf_dropout :78, lidar_max_x/y :29-32) so that a real
    ``TransfuserConfig`` instance can be passed wherever this one is accepted.
    """

    tf_d_model: int = 256
    tf_d_ffn: int = 1024
    tf_num_head: int = 8
    tf_dropout: float = 0.0
    lidar_max_x: float = 32.0
    lidar_max_y: float = 32.0
    # literals of the reference, parametrised for the stress configuration
    num_decoder_layers: int = 2      # transfuser_model_v2.py:476
    step_num: int = 2                # transfuser_model_v2.py:581
    trunc_timestep: int = 8          # transfuser_model_v2.py:594
    plan_anchor_path: str = ""


# odometry normalisation literals (transfuser_model_v2.py:485-487, 496-498)
ODO_X_OFF, ODO_X_RANGE = 1.2, 56.9
ODO_Y_OFF, ODO_Y_RANGE = 20.0, 46.0

SINE_HIDDEN = 64        # gen_sineembed_for_position(hidden_dim=64), transfuser_model_v2.py:605
NUM_TRAIN_TIMESTEPS = 1000   # DDIMScheduler ctor, transfuser_model_v2.py:447-451


def roll_timesteps(step_num: int) -> np.ndarray:
    """Denoising timesteps, transfuser_model_v2.py:585-588.

    ``(arange(S) * (20 / S)).round()[::-1]`` with numpy's round-half-even:
    S=2 -> [10, 0]; S=3 -> [13, 7, 0].
    """
    step_ratio = 20 / step_num
    return (np.arange(0, step_num) * step_ratio).round()[::-1].copy().astype(np.int64)
