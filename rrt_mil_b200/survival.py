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
    """Negative log-likelihood survival loss (Survival/utils/loss.py:25-43).  ``Y``: ground-truth bin
    ``[B]`` (long), ``c``: censorship status ``[B]`` (0 = event observed)."""
    B = len(Y)
    Y = Y.view(B, 1)
    c = c.view(B, 1).float()
    if S is None:
        S = torch.cumprod(1 - hazards, dim=1)
    S_padded = torch.cat([torch.ones_like(c), S], 1)
    uncensored = -(1 - c) * (torch.log(torch.gather(S_padded, 1, Y).clamp(min=eps)) +
                             torch.log(torch.gather(hazards, 1, Y).clamp(min=eps)))
    censored = -c * torch.log(torch.gather(S_padded, 1, Y + 1).clamp(min=eps))
    return ((1 - alpha) * (censored + uncensored) + alpha * uncensored).mean()


class SurvivalRRTMIL(RRTMIL):
    """``forward(x) -> (hazards, S)`` like the reference's Survival RRTMIL (``n_classes`` = time bins)."""

    def __init__(self, input_dim=1024, n_classes=4, **kw):
        super().__init__(input_dim=input_dim, n_classes=n_classes, **kw)

    def forward(self, x, return_attn=False, no_norm=False):
        if return_attn:
            logits, a = super().forward(x, return_attn=True, no_norm=no_norm)
            return (*hazards_and_survival(logits), a)
        return hazards_and_survival(super().forward(x))
