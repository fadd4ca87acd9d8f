"""Gradient w.r.t. the IMAGE (the drop-in modules are autograd-differentiable w.r.t. the input and all
parameters).  Training never asks for it; saliency / adversarial uses of the reference's models do.  The CUDA path (stem
transposed convolution `hno_stem_backward_input`, or conv1's input gradient for use_resize=False) against autograd through the
fp64 oracle on the fixtures recorded from the reference."""
import os

import numpy as np
import pytest
import torch

from oracle import hno_oracle as orc

pytestmark = pytest.mark.gpu


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _oracle_dx(forward, sd, x, labels, nclass):
    xd = x.double().requires_grad_(True)
    probs = forward({k: v.double() for k, v in sd.items()}, xd)
    loss = orc.dice_loss(probs, orc.to_categorical(labels, nclass).double())
    (dx,) = torch.autograd.grad(loss, xd)
    return float(loss), dx


def test_stem_module_input_gradient(cuda):
    from multimodal_3d_image_segmentation_b200 import nets
    torch.manual_seed(3)
    for cin, f, spatial in ((4, 8, (9, 10, 11)), (2, 24, (8, 7, 6)), (3, 24, (5, 5, 5))):
        conv = nets.nets_utils.ConvNormAct(cin, f, kernel_size=2, stride=2, use_bias=True, activation='selu', ndim=5, device=cuda)
        x = torch.randn(2, cin, *spatial)
        g = torch.randn(2, f, *[s // 2 + 1 for s in spatial])
        xc = x.to(cuda).requires_grad_(True)
        (conv(xc) * g.to(cuda)).sum().backward()
        xr = x.double().requires_grad_(True)
        w, b = conv.op.weight.detach().cpu().double(), conv.op.bias.detach().cpu().double()
        y = torch.nn.functional.selu(torch.nn.functional.conv3d(xr, w, b, stride=2, padding=1))
        (dx,) = torch.autograd.grad((y * g.double()).sum(), xr)
        assert rel(xc.grad, dx) < 1e-5, (cin, f, rel(xc.grad, dx))


@pytest.mark.parametrize('path', ['dropin', 'fused'])
def test_hnosegxs_input_gradient(cuda, golden_dir, path):
    from multimodal_3d_image_segmentation_b200 import nets
    g = dict(np.load(os.path.join(golden_dir, 'model_small.npz')))
    sd = _sd(g, 'shared/sd/')
    blocks, modes = [1, 2, 1, 2, 1, 2], (2, 3, 3)
    model = nets.HNOSegXS(2, 3, 8, blocks, modes, device=cuda)
    model.load_state_dict(sd)
    x = torch.from_numpy(g['shared/x'])
    labels = torch.from_numpy(g['shared/labels'].astype(np.int64))
    o_loss, o_dx = _oracle_dx(lambda p, xx: orc.hnosegxs_forward(p, xx, blocks, modes), sd, x, labels, 3)
    xc = x.to(cuda).requires_grad_(True)
    if path == 'dropin':
        loss = nets.custom_losses.DiceLoss()(model(xc), orc.to_categorical(labels, 3).to(cuda))
    else:
        loss = model.loss(xc, labels.to(cuda), 'DiceLoss')
    loss.backward()
    assert abs(float(loss) - o_loss) < 2e-6
    assert rel(xc.grad, o_dx) < 2e-4, rel(xc.grad, o_dx)
    for k, p in model.named_parameters():  # the parameter gradients are unchanged by asking for dx
        assert rel(p.grad, g[f'shared/DiceLoss/grad/{k}']) < 2e-4, k


@pytest.mark.parametrize('tag,cin', [('xs2', 2), ('xs4', 4)])
def test_noresize_input_gradient(cuda, golden_dir, tag, cin):
    from multimodal_3d_image_segmentation_b200 import nets
    g = dict(np.load(os.path.join(golden_dir, 'noresize_small.npz')))
    sd = _sd(g, f'{tag}/sd/')
    blocks, modes = [1, 2, 1, 2], (2, 3, 3)
    model = nets.HNOSegXS(cin, 3, 8, blocks, modes, use_resize=False, device=cuda)
    model.load_state_dict(sd)
    x = torch.from_numpy(g[f'{tag}/x'])
    labels = torch.from_numpy(g[f'{tag}/labels'].astype(np.int64))
    o_loss, o_dx = _oracle_dx(lambda p, xx: orc.hnosegxs_forward(p, xx, blocks, modes, use_resize=False), sd, x, labels, 3)
    xc = x.to(cuda).requires_grad_(True)
    loss = nets.custom_losses.DiceLoss()(model(xc), orc.to_categorical(labels, 3).to(cuda))
    loss.backward()
    assert abs(float(loss) - o_loss) < 2e-6
    assert rel(xc.grad, o_dx) < 2e-4, rel(xc.grad, o_dx)


def test_neural_operator_seg_input_gradient(cuda, golden_dir):
    from multimodal_3d_image_segmentation_b200 import nets
    g = dict(np.load(os.path.join(golden_dir, 'hnoseg_small.npz')))
    sd = _sd(g, 'sd/')
    model = nets.NeuralOperatorSeg(2, 3, 8, 3, (2, 3, 3), 'Hartley', device=cuda)
    model.load_state_dict(sd)
    x = torch.from_numpy(g['x'])
    labels = torch.from_numpy(g['labels'].astype(np.int64))
    o_loss, o_dx = _oracle_dx(lambda p, xx: orc.hnoseg_forward(p, xx, 3, (2, 3, 3)), sd, x, labels, 3)
    xc = x.to(cuda).requires_grad_(True)
    loss = nets.custom_losses.DiceLoss()(model(xc), orc.to_categorical(labels, 3).to(cuda))
    loss.backward()
    assert abs(float(loss) - o_loss) < 2e-6
    assert rel(xc.grad, o_dx) < 2e-4, rel(xc.grad, o_dx)
