"""Losses on probabilities, one fused reduction pass each (reference: nets/custom_losses.py:17-133)."""
import torch
from torch.nn import Module

from .. import ops


class _CudaLoss(Module):
    kind = None

    def forward(self, y_pred, y_true):
        return ops.ProbabilityLoss.apply(y_pred, y_true, self.kind)


class PCCLoss(_CudaLoss):
    """1 - (r + 1) / 2 with Pearson's r per (sample, label) over the voxels, then the mean."""
    kind = ops.LOSS_KINDS['PCCLoss']


class DiceLoss(_CudaLoss):
    """mean(1 - 2 sum(t p) / (sum(t + p) + 1e-7))."""
    kind = ops.LOSS_KINDS['DiceLoss']


class ExpDiceLoss(Module):
    """Exponential-logarithmic Dice (reference :114-133).  Not on the HNOSeg-XS hot path and not ported."""

    def __init__(self, exp=0.3):
        super().__init__()
        self.exp = exp

    def forward(self, y_pred, y_true):
        raise NotImplementedError('ExpDiceLoss is outside the CUDA hot path of hno_b200 (use DiceLoss / PCCLoss)')
