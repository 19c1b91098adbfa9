"""GPU parity of the stand-alone `_C` replacements behind the reference's layers: the sigmoid focal
loss on [n, C] logits (`paa_sigmoid_focal_loss_forward/backward`, csrc/SigmoidFocalLoss.h:10-41) against the
oracle's restatement of the reference formula (sigmoid_focal_loss.py:40-52) and its autograd gradient."""
import numpy as np
import pytest
import torch

from oracle import paa_oracle
from tests.helpers import assert_grads_close

pytestmark = pytest.mark.gpu


def _case(seed, n, C, ignore=True):
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(n, C, generator=g) * 3.0 - 2.0).clamp(-12.0, 12.0)
    targets = torch.randint(0, C + 1, (n,), generator=g, dtype=torch.int32)         # 0 = background
    if ignore:
        targets[torch.rand(n, generator=g) < 0.1] = -1                                  # ignored rows
    return logits, targets


@pytest.mark.parametrize("gamma,alpha", [(2.0, 0.25), (1.5, 0.4)])
@pytest.mark.parametrize("n,C", [(1, 1), (37, 80), (4099, 80), (257, 3)])
def test_sigmoid_focal_loss_forward_backward_against_oracle(gamma, alpha, n, C):
    from paa_b200.layers import SigmoidFocalLoss, sigmoid_focal_loss_cuda
    logits, targets = _case(1000 * n + C, n, C)
    w = torch.rand(n, C, generator=torch.Generator().manual_seed(5)) + 0.5              # upstream gradient
    # The reference formula (sigmoid_focal_loss.py:40-52) evaluated in float64 is the exact value.  In float32 --
    # how the reference's CPU path runs it -- log(1 - p) is taken from a rounded p and is off by up to 1e-2 absolute
    # for |logit| >~ 8 (11 of 328 000 elements of the largest case); the reference's CUDA kernel, which is what the
    # `_C` entry points replaced here stand for, uses the stable softplus form (SigmoidFocalLoss_cuda.cu:43-51).
    # So: everything against float64 to 1e-4, and against the float32 oracle wherever that is itself accurate.
    ref64_in = logits.double().requires_grad_(True)
    ref64 = paa_oracle.focal_loss_cpu(ref64_in, targets, gamma, alpha)
    (ref64 * w.double()).sum().backward()
    ref_in = logits.clone().requires_grad_(True)
    ref = paa_oracle.focal_loss_cpu(ref_in, targets, gamma, alpha)
    (ref * w).sum().backward()
    x = logits.cuda().requires_grad_(True)
    got = sigmoid_focal_loss_cuda(x, targets.cuda(), gamma, alpha)
    assert got.shape == (n, C)
    (got * w.cuda()).sum().backward()
    fwd, bwd = got.detach().cpu().numpy(), x.grad.cpu().numpy()
    assert_grads_close(fwd, ref64.detach().numpy(), rtol=1e-4, atol=1e-12, what="focal forward vs float64")
    assert_grads_close(bwd, ref64_in.grad.numpy(), rtol=1e-4, atol=1e-12, what="focal backward vs float64")
    ok_f = np.abs(ref.detach().numpy() - ref64.detach().numpy()) <= 2e-5 * np.abs(ref64.detach().numpy()) + 1e-12
    ok_b = np.abs(ref_in.grad.numpy() - ref64_in.grad.numpy()) <= 2e-5 * np.abs(ref64_in.grad.numpy()) + 1e-12
    assert ok_f.mean() > 0.99 and ok_b.mean() > 0.99
    assert_grads_close(np.where(ok_f, fwd, 0.0), np.where(ok_f, ref.detach().numpy(), 0.0), rtol=1e-4, atol=1e-12,
                       what="focal forward vs float32 oracle")
    assert_grads_close(np.where(ok_b, bwd, 0.0), np.where(ok_b, ref_in.grad.numpy(), 0.0), rtol=1e-4, atol=1e-12,
                       what="focal backward vs float32 oracle")
    # ignored rows and all-negative rows
    ign = (targets < 0).numpy()
    assert not got.detach().cpu().numpy()[ign].any() and not x.grad.cpu().numpy()[ign].any()
    # the module: (gamma, alpha) may come as the 1-tuples the reference's CPU formula indexes, sum=True by default
    layer = SigmoidFocalLoss((gamma,), (alpha,))
    total = layer(logits.cuda(), targets.cuda())
    np.testing.assert_allclose(float(total), float(ref.detach().double().sum()), rtol=1e-5)
    assert layer(logits.cuda(), targets.cuda(), sum=False).shape == (n, C)
    assert "gamma" in repr(layer)


def test_sigmoid_focal_loss_int64_targets_non_contiguous_logits_and_empty_input():
    from paa_b200.layers import sigmoid_focal_loss_cuda
    logits, targets = _case(9, 64, 80)
    wide = torch.zeros(64, 160)
    wide[:, ::2] = logits
    got = sigmoid_focal_loss_cuda(wide.cuda()[:, ::2], targets.long().cuda(), 2.0, 0.25)
    ref = paa_oracle.focal_loss_cpu(logits, targets, 2.0, 0.25)
    assert_grads_close(got.cpu().numpy(), ref.numpy(), rtol=1e-4, atol=1e-12, what="strided logits")
    empty = sigmoid_focal_loss_cuda(torch.zeros(0, 80, device="cuda"), torch.zeros(0, dtype=torch.int32, device="cuda"),
                                    2.0, 0.25)
    assert empty.shape == (0, 80)
    with pytest.raises(RuntimeError):        # no CPU path
        sigmoid_focal_loss_cuda(logits, targets, 2.0, 0.25)


def test_extreme_logits_stay_finite():
    """The reference's CUDA kernel clamps p at FLT_MIN (SigmoidFocalLoss_cuda.cu:43-51) where its CPU formula gives
    inf / NaN: the stable form here stays finite and keeps the right limits."""
    from paa_b200.layers import sigmoid_focal_loss_cuda
    x = torch.tensor([[-100.0, 100.0, -30.0, 30.0]], device="cuda", requires_grad=True)
    t = torch.tensor([2], dtype=torch.int32, device="cuda")          # column 1 positive
    out = sigmoid_focal_loss_cuda(x, t, 2.0, 0.25)
    out.sum().backward()
    o = out.detach().cpu().numpy()[0]
    assert np.isfinite(o).all() and np.isfinite(x.grad.cpu().numpy()).all()
    assert o[0] == 0.0 and o[1] < 1e-30                                # confident and right
    np.testing.assert_allclose(o[3], 0.75 * 30.0, rtol=1e-6)            # confident and wrong: (1-a) * p^2 * softplus(30)
