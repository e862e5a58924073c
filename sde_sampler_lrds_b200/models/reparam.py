"""Control parametrisations with the reference's names (sde_sampler/models/reparam.py: ClippedCtrl 18-43 =
``base_zero_init``, ScoreCtrl 67-117 = ``target_informed_zero_init``).  ``forward(t, x)`` runs lrds_ctrl_forward."""
from __future__ import annotations

from typing import Callable

import torch
from torch.nn import Module


class ClippedCtrl(Module):
    """clip(base_model(t, x), +-clip_model)."""

    def __init__(self, base_model: Module, clip_model: float | None = None, name: str = "ctrl", **kwargs):
        super().__init__()
        self.base_model = base_model
        self.clip_model = clip_model
        self.name = name

    def forward(self, t: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        from ..pack import ctrl_forward
        return ctrl_forward(self, t, x)


class ScoreCtrl(ClippedCtrl):
    """clip(base_model(t, x)) + scale_score * clip(target_score(x)) * clip(score_model(t))  - the target-informed
    drift model.  ``target_score`` must be the bound ``score`` of a kernel-backed Distribution."""

    def __init__(self, *args, target_score: Callable, score_model: Module | None = None, detach_score: bool = True,
                 scale_score: float = 1.0, clip_score: float | None = None, **kwargs):
        super().__init__(*args, **kwargs)
        self.score_model = score_model
        self.target_score = target_score
        self.detach_score = detach_score
        self.scale_score = scale_score
        self.clip_score = clip_score
