"""Host-side contract of the drop-in modules (no GPU): constructor signatures, state_dict layout, meta forward."""
import copy
import os

import numpy as np
import pytest
import torch

from multimodal_3d_image_segmentation_b200 import nets


def test_state_dict_layout_matches_reference_fixture(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, 'model_full_probe.npz')))
    ref = {k[3:]: v.shape for k, v in g.items() if k.startswith('sd/')}
    model = nets.HNOSegXS(4, 4, 24, [3] * 8, (10, 14, 14))
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref[k]), k
    assert sum(p.numel() for p in model.parameters()) == 28248  # reference README.md:57-63
    model.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('sd/')})
    assert model.in_channels == 4 and model.out_channels == 4


def test_small_and_individual_layouts(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, 'model_small.npz')))
    for wt in ('shared', 'individual'):
        ref = {k[len(wt) + 4:]: v.shape for k, v in g.items() if k.startswith(f'{wt}/sd/')}
        model = nets.HNOSegXS(2, 3, 8, [1, 2, 1, 2, 1, 2], (2, 3, 3), weights_type=wt)
        sd = model.state_dict()
        assert list(sd.keys()) == list(ref.keys())
        assert all(tuple(sd[k].shape) == tuple(ref[k]) for k in sd)


def test_meta_forward_and_deepcopy():
    model = nets.HNOSegXS(4, 4, 24, [3] * 8, (10, 14, 14))
    clone = copy.deepcopy(model).to('meta')  # experiments/train_test.py:118-122 (torchinfo on device='meta')
    out = clone(torch.empty(1, 4, 240, 240, 155, device='meta'))
    assert tuple(out.shape) == (1, 4, 240, 240, 155)
    assert set(dict(clone.named_parameters())) == set(dict(model.named_parameters()))


def test_unsupported_options_raise_clearly():
    with pytest.raises(NotImplementedError):
        nets.HNOSegXS(4, 4, 24, [3] * 8, (10, 14), ndim=4)
    with pytest.raises(NotImplementedError):
        nets.HNOSegXS(4, 4, 24, [3] * 8, (10, 14, 14), activation='relu')
    with pytest.raises(ValueError):
        nets.HartleyOperator(8, 8, (2, 3, 3), weights_type='bogus')
    model = nets.HNOSegXS(4, 4, 24, [3] * 8, (10, 14, 14))
    with pytest.raises(RuntimeError):  # no CPU path
        model(torch.zeros(1, 4, 16, 16, 16))


def test_padcrop_semantics():
    x = torch.arange(2 * 5 * 4 * 3, dtype=torch.float32).reshape(1, 2, 5, 4, 3)
    y = nets.spatial_padcrop(x, (4, 7, 3))
    assert tuple(y.shape) == (1, 2, 4, 7, 3)
    assert torch.equal(y[:, :, :, 1:5, :], x[:, :, 0:4])  # crop: odd voxel removed at the top; pad: 1 low, 2 high
    assert float(y[:, :, :, 0].abs().sum()) == 0 and float(y[:, :, :, 5:].abs().sum()) == 0
    assert nets.spatial_padcrop(x, (5, 4, 3)) is x


@pytest.mark.parametrize('name, kw', [
    ('hnoseg_small', dict(transform_type='Hartley')),                                   # config_hnoseg.ini
    ('hnoseg_individual_small', dict(transform_type='Hartley', weights_type='individual')),
    ('fnoseg_small', dict(transform_type='Fourier')),                                   # config_fnoseg.ini
    ('fno_small', dict(transform_type='Fourier', weights_type='individual', use_bias_conv_branch=True,
                       use_block_skip=False)),                                          # config_fno.ini
])
def test_neural_operator_seg_layouts_and_meta_forward(golden_dir, name, kw):
    """state_dict keys / shapes of the NeuralOperatorSeg family equal those recorded from the reference; meta forward."""
    g = dict(np.load(os.path.join(golden_dir, name + '.npz')))
    ref = {k[3:]: tuple(v.shape) for k, v in g.items() if k.startswith('sd/')}
    model = nets.NeuralOperatorSeg(2, 3, 8, 3, (2, 3, 3), **kw)
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == ref
    model.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('sd/')})
    meta = nets.NeuralOperatorSeg(2, 3, 8, 3, (2, 3, 3), device='meta', **kw)
    assert tuple(meta(torch.empty(1, 2, 18, 16, 13, device='meta')).shape) == (1, 3, 18, 16, 13)


def test_losses_and_input_side_host_contract():
    """Constructor arguments of the loss classes (run.py:105-110 passes the ini's [loss] section as kwargs) and the
    no-CPU-path rule of the experiments.utils mirror."""
    from multimodal_3d_image_segmentation_b200.experiments import normalize_modalities, to_categorical
    assert nets.custom_losses.ExpDiceLoss().exp == 0.3 and nets.custom_losses.ExpDiceLoss(exp=0.5).param == 0.5
    with pytest.raises(ValueError):
        nets.custom_losses.ExpDiceLoss(exp=-1.0)
    nets.custom_losses.CrossEntropyLoss()
    with pytest.raises(NotImplementedError):
        nets.custom_losses.CrossEntropyLoss(weight=torch.ones(4))
    for name in ('DiceLoss', 'PCCLoss', 'ExpDiceLoss', 'CrossEntropyLoss'):
        assert hasattr(nets.custom_losses, name)  # hasattr(custom_losses, loss_name) decides the route in run.py:107
    with pytest.raises(RuntimeError):
        to_categorical(torch.zeros(1, 1, 2, 2, 2, dtype=torch.uint8), 2)
    with pytest.raises(RuntimeError):
        normalize_modalities(torch.zeros(4, 2, 2, 2), mask_val=0)
    with pytest.raises(RuntimeError):
        nets.custom_losses.CrossEntropyLoss()(torch.zeros(1, 2, 2, 2, 2), torch.zeros(1, 2, 2, 2, 2))
