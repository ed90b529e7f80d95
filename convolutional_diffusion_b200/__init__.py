"""B200-native analytic score machines (LS / ELS / bbELS) -- drop-in for the hot path of
henhen724/convolutional_diffusion (`src/utils/idealscore.py`).  CUDA (sm_100a) only."""
from .modules import (LocalScoreModule, LocalEquivScoreModule, LocalEquivBordersScoreModule, IdealScoreModule,  # noqa: F401
                      cosine_noise_schedule, exponential_schedule, linear_noise_schedule)
from .machine import ScheduledScoreMachine, ddim_coefficients  # noqa: F401
from .bank import PatchBank  # noqa: F401
from .engine import ScoreEngine  # noqa: F401

__all__ = ["LocalScoreModule", "LocalEquivScoreModule", "LocalEquivBordersScoreModule", "IdealScoreModule",
           "ScheduledScoreMachine",
           "PatchBank", "ScoreEngine", "cosine_noise_schedule", "exponential_schedule", "linear_noise_schedule",
           "ddim_coefficients"]
