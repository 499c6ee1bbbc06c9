"""Survival variant of the model (Survival/models/RRTMIL/network.py:789-793,
Survival/utils/loss.py:25-43): the same RRTMIL trunk, with the ``n_classes`` logits read as discrete-time
hazards.  The head and the loss act on ``[1, n_bins]`` numbers, so they are plain torch on the logits the
CUDA path produced; gradients flow back into the CUDA backward through ``d loss / d logits``."""
from __future__ import annotations

import torch

from .mil import RRTMIL


def hazards_and_survival(logits: torch.Tensor):
    """``hazards = sigmoid(logits)``, ``S = cumprod(1 - hazards)`` (network.py:791-792)."""
    hazards = torch.sigmoid(logits)
    return hazards, torch.cumprod(1 - hazards, dim=1)


def nll_surv_loss(hazards, S, Y, c, alpha: float = 0.0, eps: float = 1e-7):
    """Discrete-time survival negative log-likelihood, same value as the reference's ``nll_loss``
    (Survival/utils/loss.py:25-43), stated per sample instead of through a padded survival table:

        event observed in bin y (c = 0):   -( log S(y - 1) + log h(y) ),   S(-1) = 1
        censored in bin y       (c = 1):   -  log S(y)

    and ``loss = mean(event + (1 - alpha) * censored)``.  ``hazards``, ``S``: ``[B, n_bins]``; ``Y``: bin index
    ``[B]``; ``c``: censorship flag ``[B]``.  Logs are taken of values clamped at ``eps``."""
    y = Y.reshape(-1).long()
    cens = c.reshape(-1).to(hazards.dtype)
    if S is None:
        S = (1 - hazards).cumprod(dim=1)
    rows = torch.arange(y.numel(), device=y.device)
    log_s_here = S[rows, y].clamp_min(eps).log()
    s_before = torch.where(y > 0, S[rows, (y - 1).clamp_min(0)], torch.ones_like(log_s_here))
    log_event = s_before.clamp_min(eps).log() + hazards[rows, y].clamp_min(eps).log()
    per_sample = -(1 - cens) * log_event - (1 - alpha) * cens * log_s_here
    return per_sample.mean()


class SurvivalRRTMIL(RRTMIL):
    """``forward(x) -> (hazards, S)`` like the reference's Survival RRTMIL (``n_classes`` = time bins)."""

    def __init__(self, input_dim=1024, n_classes=4, **kw):
        super().__init__(input_dim=input_dim, n_classes=n_classes, **kw)

    def forward(self, x, return_attn=False, no_norm=False):
        if return_attn:
            logits, a = super().forward(x, return_attn=True, no_norm=no_norm)
            return (*hazards_and_survival(logits), a)
        return hazards_and_survival(super().forward(x))
