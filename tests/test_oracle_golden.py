"""The oracle restatement replayed against fixtures produced by the real reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import hno_oracle as orc


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + '.npz')))


def _close(a, b, tol=2e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    assert err < tol, err


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def test_dht_fixtures(golden_dir):
    g = _load(golden_dir, 'dht')
    x = torch.from_numpy(g['x'])
    _close(orc.dhtn(x), g['fwd'])
    _close(orc.dhtn(x, inverse=True), g['inv'])
    for tag in 'abc':
        xx = torch.from_numpy(g[f'x_{tag}'])
        modes = tuple(int(v) for v in g[f'modes_{tag}'])
        z = orc.transform_crop(xx, modes)
        _close(z, g[f'z_{tag}'])
        _close(orc.pad_inverse(torch.from_numpy(g[f'z_{tag}']), xx.shape[2:]), g[f'y_{tag}'])
        m = orc.clamp_modes(modes, xx.shape[2:])
        kl = [orc.corner_indices(n, mm) for n, mm in zip(xx.shape[2:], m)]
        _close(orc.dht3_dense(g[f'x_{tag}'], kl, 1.0 / np.prod(xx.shape[2:])), g[f'z_{tag}'], 1e-6)


def test_operator_fixtures(golden_dir):
    g = _load(golden_dir, 'operator')
    z = torch.from_numpy(g['z'])
    for wt in ('shared', 'individual'):
        w = torch.from_numpy(g[f'w_{wt}']).requires_grad_(True)
        zz = z.clone().requires_grad_(True)
        y = orc.hartley_mix(zz, w)
        _close(y.detach(), g[f'y_{wt}'])
        dz, dw = torch.autograd.grad(y, [zz, w], torch.from_numpy(g[f'g_{wt}']))
        _close(dz, g[f'dz_{wt}'])
        _close(dw, g[f'dw_{wt}'])
    _close(orc.hartley_operator_with_transform(torch.from_numpy(g['x_t']), torch.from_numpy(g['w_t']), (2, 3, 3)),
           g['y_t'])


def test_block_fixtures(golden_dir):
    g = _load(golden_dir, 'block')
    for tag in ('plain', 'mapped'):
        sd = _sd(g, f'sd_{tag}/')
        y = orc.xs_block(torch.from_numpy(g[f'x_{tag}']), sd, 'layers.0.', 2, (2, 3, 3))
        _close(y, g[f'y_{tag}'])


def test_loss_fixtures(golden_dir):
    g = _load(golden_dir, 'losses')
    for name in ('DiceLoss', 'PCCLoss', 'ExpDiceLoss', 'CrossEntropyLoss'):
        p = torch.from_numpy(g['p']).requires_grad_(True)
        loss = orc.LOSSES[name](p, torch.from_numpy(g['t']))
        _close(loss.detach(), g[f'{name}/loss'], 1e-6)
        _close(torch.autograd.grad(loss, p)[0], g[f'{name}/grad'], 1e-5)


def test_small_model_fixtures(golden_dir):
    g = _load(golden_dir, 'model_small')
    blocks, modes = [1, 2, 1, 2, 1, 2], (2, 3, 3)
    for wt in ('shared', 'individual'):
        sd = _sd(g, f'{wt}/sd/')
        x = torch.from_numpy(g[f'{wt}/x'])
        probs, logits = orc.hnosegxs_forward(sd, x, blocks, modes, return_logits=True)
        _close(probs, g[f'{wt}/probs'])
        _close(logits, g[f'{wt}/logits'])
        labels = torch.from_numpy(g[f'{wt}/labels'].astype(np.int64))
        for lname in ('DiceLoss', 'PCCLoss'):
            loss, grads = orc.train_step(sd, x, labels, blocks, modes, lname)
            _close(loss, g[f'{wt}/{lname}/loss'], 1e-6)
            for k, v in grads.items():
                _close(v, g[f'{wt}/{lname}/grad/{k}'], 2e-4)


def test_param_count_known_answer():
    # the reference's only published known answer: README.md:57-63
    sd = orc.init_state_dict(4, 4, 24, [3] * 8, (10, 14, 14))
    assert sum(v.numel() for v in sd.values()) == 28248


@pytest.mark.parametrize('name', ['hnoseg_small', 'hnoseg_individual_small'])
def test_hnoseg_fixture(golden_dir, name):
    """NeuralOperatorSeg(transform_type='Hartley') = HNOSeg, shared and per-mode weights, recorded from
    the real reference."""
    g = _load(golden_dir, name)
    sd = _sd(g, 'sd/')
    x = torch.from_numpy(g['x'])
    probs, logits = orc.hnoseg_forward(sd, x, 3, (2, 3, 3), return_logits=True)
    _close(probs, g['probs'])
    _close(logits, g['logits'])
    labels = torch.from_numpy(g['labels'].astype(np.int64))
    loss, grads = orc.hnoseg_train_step(sd, x, labels, 3, (2, 3, 3), 'DiceLoss')
    _close(loss, g['DiceLoss/loss'], 1e-6)
    for k, v in grads.items():
        _close(v, g[f'DiceLoss/grad/{k}'], 2e-4)


def test_fourier_fixtures(golden_dir):
    """FourierOperator and NeuralOperatorSeg(transform_type='Fourier') = FNOSeg, recorded from the real reference."""
    g = _load(golden_dir, 'fourier_operator')
    for tag in ('a', 'b', 'c', 'd'):  # c, d: per-mode ('individual') complex weights
        y = orc.fourier_operator_with_transform(torch.from_numpy(g[f'{tag}/x']), torch.from_numpy(g[f'{tag}/wr']),
                                                torch.from_numpy(g[f'{tag}/wi']), tuple(int(v) for v in g[f'{tag}/modes']))
        _close(y, g[f'{tag}/y'])
    g = _load(golden_dir, 'fnoseg_small')
    sd = _sd(g, 'sd/')
    probs, logits = orc.hnoseg_forward(sd, torch.from_numpy(g['x']), 3, (2, 3, 3), return_logits=True)
    _close(probs, g['probs'])
    _close(logits, g['logits'])
    # experiments/config_files/config_fno.ini in small: individual weights, biased conv branch, no block skip
    g = _load(golden_dir, 'fno_small')
    sd = _sd(g, 'sd/')
    assert tuple(sd['layers.0.op.weight_real'].shape) == (8, 8, 4, 6, 3) and 'layers.0.conv_concat.op.weight' not in sd
    probs, logits = orc.hnoseg_forward(sd, torch.from_numpy(g['x']), 3, (2, 3, 3), return_logits=True,
                                       use_block_skip=False)
    _close(probs, g['probs'])
    _close(logits, g['logits'])
    loss, grads = orc.hnoseg_train_step(sd, torch.from_numpy(g['x']), torch.from_numpy(g['labels'].astype(np.int64)), 3,
                                        (2, 3, 3), 'DiceLoss', use_block_skip=False)
    _close(loss, g['DiceLoss/loss'], 1e-6)
    for k, v in grads.items():
        _close(v, g[f'DiceLoss/grad/{k}'], 2e-4)


def test_input_side_fixtures(golden_dir):
    """to_categorical / normalize_modalities restatements against outputs of the reference's experiments/utils.py."""
    g = _load(golden_dir, 'input_side')
    onehot = orc.to_categorical(torch.from_numpy(g['labels'].astype(np.int64)), 4)
    assert np.array_equal(onehot.numpy(), g['onehot'])
    cases = {'plain': {}, 'mask': dict(mask_val=0), 'clip': dict(clip_val=(50.0, 900.0)),
             'maskclip': dict(mask_val=0, clip_val=(0.0, 700.0)), 'maskhit': dict(mask_val=700, clip_val=(0.0, 700.0))}
    for tag, kw in cases.items():
        y = orc.normalize_modalities(g['vol'], **kw)
        assert y.dtype == np.float32 and np.abs(y - g[f'norm/{tag}']).max() < 1e-5, tag


def test_hartley_operator_with_transform_individual_fixtures(golden_dir):
    """HartleyOperator(use_transform=True, weights_type='individual'), recorded from the real reference."""
    g = _load(golden_dir, 'operator_transform_individual')
    for tag in ('a', 'b', 'c'):
        x = torch.from_numpy(g[f'{tag}/x']).requires_grad_(True)
        w = torch.from_numpy(g[f'{tag}/w']).requires_grad_(True)
        y = orc.hartley_operator_with_transform_individual(x, w, tuple(int(v) for v in g[f'{tag}/modes']))
        _close(y.detach(), g[f'{tag}/y'])
        dx, dw = torch.autograd.grad((y * torch.from_numpy(g[f'{tag}/g'])).sum(), [x, w])
        _close(dx, g[f'{tag}/dx'], 1e-5)
        _close(dw, g[f'{tag}/dw'], 1e-5)


def test_hartley_mha_fixtures(golden_dir):
    """HartleyMultiHeadAttention (BASELINE config 5's layer): the oracle restatement against outputs and
    gradients of the real reference -- grouped / plain self-attention, value_dim != key_dim, cross-attention.  The CUDA
    path for this layer is not built yet; the oracle is pinned first."""
    g = _load(golden_dir, 'hartley_mha')
    for tag, nin in (('self_grouped', 1), ('self_plain', 1), ('cross', 3)):
        xs = [torch.from_numpy(g[f'{tag}/x{i}']).requires_grad_(True) for i in range(nin)]
        ws = [torch.from_numpy(g[f'{tag}/{k}']).requires_grad_(True) for k in ('wq', 'wk', 'wv', 'wo')]
        patch = tuple(int(v) for v in g[f'{tag}/patch'])
        y = orc.hartley_mha(xs[0], *ws, tuple(int(v) for v in g[f'{tag}/modes']), patch if patch[0] else None,
                            key=xs[1] if nin > 1 else None, value=xs[2] if nin > 2 else None)
        _close(y.detach(), g[f'{tag}/y'])
        grads = torch.autograd.grad((y * torch.from_numpy(g[f'{tag}/g'])).sum(), xs + ws)
        for i in range(nin):
            _close(grads[i], g[f'{tag}/dx{i}'], 1e-5)
        for k, gw in zip(('dwq', 'dwk', 'dwv', 'dwo'), grads[nin:]):
            _close(gw, g[f'{tag}/{k}'], 1e-5)
    # grouping is a pure permutation and ungrouping its inverse
    t = torch.randn(2, 3, 4, 4, 6, 2)
    assert torch.equal(orc.ungroup_patches(orc.group_patches(t, (2, 3, 1)), 4, (2, 3, 1)), t)


def test_hartley_mha_seg_fixture(golden_dir):
    """HartleyMHASeg (BASELINE config 5's architecture, deep supervision on) in small: oracle vs the real reference."""
    g = _load(golden_dir, 'hartley_mha_seg_small')
    sd = _sd(g, 'sd/')
    assert 'conv_ds.op.weight' in sd and tuple(sd['conv_ds.op.weight'].shape) == (3, 24, 1, 1, 1)
    x = torch.from_numpy(g['x'])
    _close(orc.hnoseg_forward(sd, x, 2, (2, 4, 2), patch=(2, 2, 2)), g['probs'])
    loss, grads = orc.hnoseg_train_step(sd, x, torch.from_numpy(g['labels'].astype(np.int64)), 2, (2, 4, 2), 'DiceLoss',
                                        patch=(2, 2, 2))
    _close(loss, g['DiceLoss/loss'], 1e-6)
    for k, v in grads.items():
        _close(v, g[f'DiceLoss/grad/{k}'], 2e-4)
