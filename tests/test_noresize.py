"""use_resize=False: the networks at the image resolution, no stem and no interpolation (reference nets/hnosegxs.py:102-109,
150, 174-180; nets/architectures.py:286-289, 345-351).  Fixture tests/golden/noresize_small.npz is recorded from the REAL
reference by oracle/make_golden.py::case_noresize."""
import os

import numpy as np
import pytest
import torch

from oracle import hno_oracle as orc

XS_CASES = (('xs2', 2), ('xs4', 4))
XS_BLOCKS, MODES = [1, 2, 1, 2], (2, 3, 3)


def _load(golden_dir):
    return dict(np.load(os.path.join(golden_dir, 'noresize_small.npz')))


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_oracle_against_reference_fixture(golden_dir):
    g = _load(golden_dir)
    for tag, _ in XS_CASES:
        sd = _sd(g, f'{tag}/sd/')
        x = torch.from_numpy(g[f'{tag}/x'])
        labels = torch.from_numpy(g[f'{tag}/labels'].astype(np.int64))
        assert rel(orc.hnosegxs_forward(sd, x, XS_BLOCKS, MODES, use_resize=False), g[f'{tag}/probs']) < 2e-5
        loss, grads = orc.train_step(sd, x, labels, XS_BLOCKS, MODES, 'DiceLoss')
        assert abs(float(loss) - float(g[f'{tag}/loss'])) < 1e-6
        for k, v in grads.items():
            assert rel(v, g[f'{tag}/grad/{k}']) < 2e-4, k
    sd = _sd(g, 'hnoseg/sd/')
    x = torch.from_numpy(g['hnoseg/x'])
    assert rel(orc.hnoseg_forward(sd, x, 2, MODES), g['hnoseg/probs']) < 2e-5


def test_state_dict_layout_without_stem(golden_dir):
    from multimodal_3d_image_segmentation_b200 import nets
    g = _load(golden_dir)
    for tag, cin in XS_CASES:
        model = nets.HNOSegXS(cin, 3, 8, XS_BLOCKS, MODES, use_resize=False)
        ref = _sd(g, f'{tag}/sd/')
        assert list(model.state_dict().keys()) == list(ref.keys()) and model.conv_in is None
        model.load_state_dict(ref)
        y = model(torch.empty(1, cin, 6, 6, 6, device='meta'))
        assert tuple(y.shape) == (1, 3, 6, 6, 6)
    model = nets.NeuralOperatorSeg(3, 3, 8, 2, MODES, 'Hartley', use_resize=False, use_deep_supervision=True)
    ref = _sd(g, 'hnoseg/sd/')
    assert list(model.state_dict().keys()) == list(ref.keys())
    model.load_state_dict(ref)
    assert tuple(model(torch.empty(1, 3, 6, 6, 6, device='meta')).shape) == (1, 3, 6, 6, 6)


@pytest.mark.gpu
@pytest.mark.parametrize('tag,cin', XS_CASES)
def test_hnosegxs_against_reference_fixture(cuda, golden_dir, tag, cin):
    from multimodal_3d_image_segmentation_b200 import nets
    from multimodal_3d_image_segmentation_b200.parallel import Trainer
    g = _load(golden_dir)
    model = nets.HNOSegXS(cin, 3, 8, XS_BLOCKS, MODES, use_resize=False, device=cuda)
    model.load_state_dict(_sd(g, f'{tag}/sd/'))
    x = torch.from_numpy(g[f'{tag}/x']).to(cuda)
    labels = torch.from_numpy(g[f'{tag}/labels'].astype(np.int64)).to(cuda)
    with torch.no_grad():
        probs = model(x)
        logits = model.forward_logits(x)
        pred = model.predict_labels(x)
    assert rel(probs, g[f'{tag}/probs']) < 1e-5
    assert rel(torch.softmax(logits, 1), g[f'{tag}/probs']) < 1e-5
    assert torch.equal(pred.long(), logits.argmax(1))
    onehot = orc.to_categorical(labels.cpu(), 3).to(cuda)
    for path in ('dropin', 'fused'):
        model.zero_grad()
        if path == 'dropin':  # exactly experiments/train_test.py:159-170
            loss = nets.custom_losses.DiceLoss()(model(x), onehot)
        else:
            loss = model.loss(x, labels, 'DiceLoss')
        loss.backward()
        assert abs(float(loss) - float(g[f'{tag}/loss'])) < 2e-6, path
        for k, p in model.named_parameters():
            assert rel(p.grad, g[f'{tag}/grad/{k}']) < 2e-4, (path, k, rel(p.grad, g[f'{tag}/grad/{k}']))
    for use_graph in (False, True):
        tr = Trainer(model, loss_name='DiceLoss', use_graph=use_graph)
        loss = tr.loss_and_grad_graphed(x, labels) if use_graph else tr.loss_and_grad(x, labels)
        assert abs(float(loss) - float(g[f'{tag}/loss'])) < 2e-6
        for k, p in model.named_parameters():
            assert rel(tr.flat.grad_view_of(p), g[f'{tag}/grad/{k}']) < 2e-4, ('trainer', use_graph, k)


@pytest.mark.gpu
def test_neural_operator_seg_against_reference_fixture(cuda, golden_dir):
    """NeuralOperatorSeg('Hartley'), deep supervision on, 3 input channels (zero-padded to 4 on the device)."""
    from multimodal_3d_image_segmentation_b200 import nets
    g = _load(golden_dir)
    model = nets.NeuralOperatorSeg(3, 3, 8, 2, MODES, 'Hartley', use_resize=False, use_deep_supervision=True, device=cuda)
    model.load_state_dict(_sd(g, 'hnoseg/sd/'))
    x = torch.from_numpy(g['hnoseg/x']).to(cuda)
    labels = torch.from_numpy(g['hnoseg/labels'].astype(np.int64))
    onehot = orc.to_categorical(labels, 3).to(cuda)
    probs = model(x)
    assert rel(probs, g['hnoseg/probs']) < 1e-5
    loss = nets.custom_losses.DiceLoss()(probs, onehot)
    loss.backward()
    assert abs(float(loss) - float(g['hnoseg/loss'])) < 2e-6
    for k, p in model.named_parameters():
        assert rel(p.grad, g[f'hnoseg/grad/{k}']) < 2e-4, (k, rel(p.grad, g[f'hnoseg/grad/{k}']))


@pytest.mark.gpu
def test_hnosegxs_larger_grid_against_oracle(cuda):
    """24 filters on a 40 x 36 x 32 grid (the tensor-core pointwise / transform kernels engage from 4,096 voxels per
    sample): probabilities, Dice loss and every gradient against the fp64 oracle."""
    from multimodal_3d_image_segmentation_b200 import nets
    torch.manual_seed(5)
    blocks, modes, spatial = [2, 1, 2, 1], (5, 6, 6), (40, 36, 32)
    model = nets.HNOSegXS(4, 4, 24, blocks, modes, use_resize=False, device=cuda)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    x = torch.randn(2, 4, *spatial)
    labels = torch.randint(0, 4, (2, 1) + spatial)
    o_loss, o_grads = orc.train_step({k: v.double() for k, v in sd.items()}, x.double(), labels, blocks, modes, 'DiceLoss')
    with torch.no_grad():
        o_probs = orc.hnosegxs_forward(sd, x, blocks, modes, use_resize=False)
    loss = model.loss(x.to(cuda), labels.to(cuda), 'DiceLoss')
    loss.backward()
    with torch.no_grad():
        assert rel(model(x.to(cuda)), o_probs) < 1e-4
    assert abs(float(loss) - float(o_loss)) < 1e-5
    flat = torch.cat([p.grad.flatten().cpu().double() for _, p in model.named_parameters()])
    oflat = torch.cat([o_grads[k].flatten() for k, _ in model.named_parameters()])
    assert rel(flat, oflat) < 5e-4, rel(flat, oflat)
