"""The reference's focal-loss layer on libpaa_b200.so.

Mirror of paa_core/layers/sigmoid_focal_loss.py:9-76 for callers that use the layer directly (ATSS / FCOS /
RetinaNet heads in the reference construct ``SigmoidFocalLoss(gamma, alpha)`` and call it on flattened
``[n, C]`` logits with ``[n]`` integer targets: class c > 0 marks column c-1 positive, t >= 0 the other
columns negative, t < 0 the whole row ignored).  ``_C.sigmoid_focalloss_forward`` / ``_backward``
(csrc/SigmoidFocalLoss.h:10-41) become ``paa_sigmoid_focal_loss_forward`` / ``_backward`` of
include/paa_b200.h; the arithmetic follows the stable form of csrc/cuda/SigmoidFocalLoss_cuda.cu:20-101.
The fused evaluators of ``paa_b200.loss`` never call this layer (they never materialise ``[n, C]``).

There is no CPU path: a CPU tensor raises, where the reference would fall back to a formula that turns
NaN beyond |logit| ~ 17 (sigmoid_focal_loss.py:40-52).
"""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from paa_b200 import _lib
from paa_b200.config import scalar


def _check(logits, targets):
    if not logits.is_cuda or not targets.is_cuda:
        raise RuntimeError("paa_b200 has no CPU path: sigmoid focal loss needs CUDA tensors")
    if logits.dim() != 2 or targets.dim() != 1 or targets.shape[0] != logits.shape[0]:
        raise RuntimeError("logits must be [n, C] and targets [n], got %s and %s"
                           % (tuple(logits.shape), tuple(targets.shape)))
    if logits.dtype != torch.float32:
        raise RuntimeError("logits must be float32, got %s" % (logits.dtype,))


class _SigmoidFocalLoss(Function):
    @staticmethod
    def forward(ctx, logits, targets, gamma, alpha):
        _check(logits, targets)
        lib = _lib.load()
        logits_c = logits.contiguous()
        targets_c = targets.to(torch.int32).contiguous()          # SigmoidFocalLoss_cuda.cu reads int targets
        ctx.save_for_backward(logits_c, targets_c)
        ctx.gamma, ctx.alpha = scalar(gamma), scalar(alpha)
        n, num_classes = logits_c.shape
        losses = torch.empty_like(logits_c)
        with _lib.device_guard(logits.device):
            _lib.check(lib.paa_sigmoid_focal_loss_forward(logits_c.data_ptr(), targets_c.data_ptr(), n, num_classes,
                                                          ctx.gamma, ctx.alpha, losses.data_ptr(),
                                                          _lib.stream_handle(logits.device)),
                       "paa_sigmoid_focal_loss_forward")
        return losses

    @staticmethod
    @once_differentiable
    def backward(ctx, d_loss):
        logits, targets = ctx.saved_tensors
        lib = _lib.load()
        d_loss = d_loss.contiguous()
        n, num_classes = logits.shape
        d_logits = torch.empty_like(logits)
        with _lib.device_guard(logits.device):
            _lib.check(lib.paa_sigmoid_focal_loss_backward(logits.data_ptr(), targets.data_ptr(), d_loss.data_ptr(), n,
                                                           num_classes, ctx.gamma, ctx.alpha, d_logits.data_ptr(),
                                                           _lib.stream_handle(logits.device)),
                       "paa_sigmoid_focal_loss_backward")
        return d_logits, None, None, None


sigmoid_focal_loss_cuda = _SigmoidFocalLoss.apply


class SigmoidFocalLoss(nn.Module):
    """sigmoid_focal_loss.py:55-76: ``forward(logits, targets, sum=True)``."""

    def __init__(self, gamma, alpha):
        super(SigmoidFocalLoss, self).__init__()
        self.gamma = gamma
        self.alpha = alpha

    def forward(self, logits, targets, sum=True):
        loss = sigmoid_focal_loss_cuda(logits, targets, self.gamma, self.alpha)
        return loss.sum() if sum == True else loss    # noqa: E712 - the reference's own test

    def __repr__(self):
        return self.__class__.__name__ + "(gamma=" + str(self.gamma) + ", alpha=" + str(self.alpha) + ")"
