"""Control parametrisations with the reference's names (sde_sampler/models/reparam.py: ClippedCtrl 18-43 =
``base_zero_init``, ScoreCtrl 67-117 = ``target_informed_zero_init``, CancelDriftCtrl 120-147 =
``target_informed_langevin_init``, LerpCtrl 150-199 = ``target_informed_lerp_tempering``).  ``forward(t, x)`` runs
lrds_ctrl_forward."""
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


class CancelDriftCtrl(ScoreCtrl):
    """Langevin initialisation (DIS): clip(base) + sde.drift(t, x) / sde.diff(t) + 0.5 sde.diff(t) score."""

    def __init__(self, *args, sde, langevin_init: bool = True, use_rescaling=True, **kwargs):
        super().__init__(*args, **kwargs)
        if sde.noise_type not in ["diagonal", "scalar"]:
            raise ValueError(f"Invalid sde noise type {sde.noise_type}.")
        if not use_rescaling:
            raise NotImplementedError("use_rescaling=False is unreachable from the shipped configs (SURVEY.md App. B.4)")
        self.sde = sde
        self.langevin_init = langevin_init
        self.use_rescaling = use_rescaling


class LerpCtrl(ScoreCtrl):
    """clip(base) + sde.diff(t) * scale_score * clip(lerp(prior_score(x), target_score(x), t / T)) * clip(score_model(t)).
    ``prior_score`` must be the bound ``score`` of the (diagonal Gaussian) prior the rollout starts from."""

    def __init__(self, *args, sde, prior_score: Callable, hard_constrain: bool = False, scale_lerp: float = 1.0, **kwargs):
        super().__init__(*args, **kwargs)
        if sde.noise_type not in ["diagonal", "scalar"]:
            raise ValueError(f"Invalid sde noise type {sde.noise_type}.")
        if hard_constrain:
            raise NotImplementedError("hard_constrain=True is not set by any shipped config (conf/model/lerp.yaml)")
        self.sde = sde
        self.prior_score = prior_score
        self.hard_constrain = hard_constrain
        self.scale_lerp = scale_lerp
