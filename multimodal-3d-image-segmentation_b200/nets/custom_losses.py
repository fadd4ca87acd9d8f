"""Losses on probabilities, one fused reduction pass each (reference: nets/custom_losses.py:17-133, plus the
torch.nn.CrossEntropyLoss fall-through of experiments/run.py:105-110)."""
import torch
from torch.nn import Module

from .. import ops


class _CudaLoss(Module):
    kind = None
    param = 0.0

    def forward(self, y_pred, y_true):
        return ops.ProbabilityLoss.apply(y_pred, y_true, self.kind, self.param)


class PCCLoss(_CudaLoss):
    """1 - (r + 1) / 2 with Pearson's r per (sample, label) over the voxels, then the mean."""
    kind = ops.LOSS_KINDS['PCCLoss']


class DiceLoss(_CudaLoss):
    """mean(1 - 2 sum(t p) / (sum(t + p) + 1e-7))."""
    kind = ops.LOSS_KINDS['DiceLoss']


class ExpDiceLoss(_CudaLoss):
    """mean((-log(clamp(dice, 1e-7, 1 - 1e-7))) ** exp) (reference :114-133): the same five-moment pass as DiceLoss,
    the exponent is applied per (sample, label) in the finalize kernel."""
    kind = ops.LOSS_KINDS['ExpDiceLoss']

    def __init__(self, exp=0.3):
        super().__init__()
        if not exp > 0:
            raise ValueError('ExpDiceLoss needs exp > 0')
        self.exp = exp
        self.param = float(exp)


class CrossEntropyLoss(Module):
    """`loss_name = CrossEntropyLoss` of the reference's ini files.  The reference has no such class in custom_losses and
    falls through to torch.nn.CrossEntropyLoss() (experiments/run.py:105-110), which it then calls on the network's
    softmax output and one-hot float targets (train_test.py:159-160).  This class keeps exactly that arithmetic
    (log-softmax OF the probabilities, class-probability targets, mean over batch x voxels) in one CUDA pass; only the
    default constructor arguments of torch.nn.CrossEntropyLoss are supported."""

    def __init__(self, weight=None, ignore_index=-100, reduction='mean', label_smoothing=0.0):
        super().__init__()
        if weight is not None or reduction != 'mean' or label_smoothing != 0.0:
            raise NotImplementedError("hno_b200 CrossEntropyLoss supports weight=None, reduction='mean', "
                                      'label_smoothing=0 only')
        # ignore_index has no effect on class-probability targets (torch ignores it there as well)

    def forward(self, y_pred, y_true):
        return ops.CrossEntropyOnProbabilities.apply(y_pred, y_true)
