"""Pins oracle/hno_oracle.py against the REAL reference and writes the fixtures under tests/golden/.

Run in the build container only (it imports /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py [--full]

For every case it (1) runs the reference module, (2) runs the oracle restatement on the same inputs and
parameters, (3) asserts they agree to fp32 round-off, (4) stores inputs + reference outputs as a small
.npz.  tests/test_oracle_golden.py replays the fixtures against the oracle; the GPU tests replay them
against the CUDA path.  --full additionally runs the BASELINE-size model (1x4x240x240x155) and stores the
reference logits at 8192 sampled voxels.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

import nets as ref  # noqa: E402  (the reference package)
from nets import hnosegxs as ref_xs  # noqa: E402
from nets.hartley_operator import HartleyOperator  # noqa: E402
from nets import custom_losses as ref_losses  # noqa: E402
from nets.dht import dhtn as ref_dhtn  # noqa: E402

from oracle import hno_oracle as orc  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def np_sd(module):
    return {k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


def check(name, a, b, tol=2e-5):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
    print(f'  {name:42s} max-rel-err {err:.2e}')
    assert err < tol, (name, err)


def save(name, **arrays):
    path = os.path.join(GOLDEN, name + '.npz')
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f'wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)')


def case_dht():
    torch.manual_seed(11)
    x = torch.randn(2, 3, 9, 8, 7)
    fwd = ref_dhtn(x, dim=(-3, -2, -1))
    inv = ref_dhtn(x, dim=(-3, -2, -1), is_inverse=True)
    check('dhtn forward', orc.dhtn(x), fwd)
    check('dhtn inverse', orc.dhtn(x, inverse=True), inv)
    kl = [list(range(9)), list(range(8)), list(range(7))]
    check('dense cas definition', orc.dht3_dense(x.numpy(), kl, 1.0 / (9 * 8 * 7)), fwd, 1e-6)
    out = {'x': x.numpy(), 'fwd': fwd.numpy(), 'inv': inv.numpy()}
    for tag, shape, modes in (('a', (1, 2, 9, 8, 7), (2, 3, 3)), ('b', (2, 2, 12, 10, 9), (10, 14, 14)),
                              ('c', (1, 3, 16, 11, 10), (4, 3, 5))):
        xx = torch.randn(*shape)
        tc = ref_xs.TransformCrop(modes, 5)
        z = tc(xx)
        check(f'TransformCrop {shape} {modes}', orc.transform_crop(xx, modes), z)
        m = orc.clamp_modes(modes, shape[2:])
        kl = [orc.corner_indices(n, mm) for n, mm in zip(shape[2:], m)]
        check(f'TransformCrop dense {tag}', orc.dht3_dense(xx.numpy(), kl, 1.0 / np.prod(shape[2:])), z, 1e-6)
        pi = ref_xs.PadInverse(5)
        y = pi(z, shape[2:])
        check(f'PadInverse {shape}', orc.pad_inverse(z, shape[2:]), y)
        out.update({f'x_{tag}': xx.numpy(), f'z_{tag}': z.numpy(), f'y_{tag}': y.numpy(),
                    f'modes_{tag}': np.array(modes)})
    save('dht', **out)


def case_operator():
    torch.manual_seed(12)
    out = {}
    z = torch.randn(2, 8, 4, 6, 6)
    for wt in ('shared', 'individual'):
        op = HartleyOperator(8, 8, (2, 3, 3), weights_type=wt, use_transform=False)
        torch.nn.init.normal_(op.weight, std=0.3)
        y = op(z)
        check(f'HartleyOperator notransform {wt}', orc.hartley_mix(z, op.weight.detach()), y.detach())
        g = torch.randn_like(y)
        zz = z.clone().requires_grad_(True)
        gy = torch.autograd.grad(op(zz), [zz, op.weight], g)
        out.update({f'w_{wt}': op.weight.detach().numpy(), f'y_{wt}': y.detach().numpy(), f'g_{wt}': g.numpy(),
                    f'dz_{wt}': gy[0].numpy(), f'dw_{wt}': gy[1].numpy()})
    out['z'] = z.numpy()
    x = torch.randn(1, 8, 9, 8, 7)
    op = HartleyOperator(8, 8, (2, 3, 3), weights_type='shared', use_transform=True)
    torch.nn.init.normal_(op.weight, std=0.3)
    y = op(x)
    check('HartleyOperator with transform', orc.hartley_operator_with_transform(x, op.weight.detach(), (2, 3, 3)),
          y.detach())
    out.update({'x_t': x.numpy(), 'w_t': op.weight.detach().numpy(), 'y_t': y.detach().numpy()})
    save('operator', **out)


def case_operator_transform_individual():
    """HartleyOperator(use_transform=True, weights_type='individual') (hartley_operator.py:196-241): forward + gradients,
    incl. an axis that the two corners fill exactly (n == 2m)."""
    out = {}
    for tag, shape, modes in (('a', (9, 8, 7), (2, 3, 3)), ('b', (9, 8, 7), (2, 4, 3)), ('c', (6, 11, 8), (3, 2, 4))):
        torch.manual_seed(14)
        op = HartleyOperator(8, 8, modes, weights_type='individual', use_transform=True)
        torch.nn.init.normal_(op.weight, std=0.3)
        x = torch.randn(2, 8, *shape, requires_grad=True)
        y = op(x)
        g = torch.randn(y.shape)
        (y * g).sum().backward()
        check(f'HartleyOperator with transform, individual {tag}',
              orc.hartley_operator_with_transform_individual(x.detach(), op.weight.detach(), modes), y.detach())
        out.update({f'{tag}/x': x.detach().numpy(), f'{tag}/w': op.weight.detach().numpy(), f'{tag}/y': y.detach().numpy(),
                    f'{tag}/g': g.numpy(), f'{tag}/dx': x.grad.numpy(), f'{tag}/dw': op.weight.grad.numpy(),
                    f'{tag}/modes': np.array(modes)})
    save('operator_transform_individual', **out)


def case_hartley_mha():
    """HartleyMultiHeadAttention (BASELINE config 5's layer): self-attention with and without patch
    grouping, value_dim != key_dim, and cross-attention with separate key / value inputs; forward + gradients."""
    from nets.hartley_mha import HartleyMultiHeadAttention
    out = {}
    cases = (('self_grouped', dict(in_channels=8, key_dim=6, num_heads=2, num_modes=(2, 4, 2), patch_size=(2, 2, 2)), 1),
             ('self_plain', dict(in_channels=8, key_dim=5, num_heads=3, num_modes=(2, 3, 3), value_dim=4), 1),
             ('cross', dict(in_channels=8, key_dim=4, num_heads=2, num_modes=(2, 2, 3), patch_size=(1, 2, 3),
                            key_in_channels=6, value_in_channels=5), 3))
    for tag, kw, nin in cases:
        torch.manual_seed(51)
        op = HartleyMultiHeadAttention(**kw)
        for w in (op.weight_query, op.weight_key, op.weight_value, op.weight_out):
            torch.nn.init.normal_(w, std=0.4)
        chans = [kw['in_channels'], kw.get('key_in_channels', kw['in_channels']),
                 kw.get('value_in_channels', kw.get('key_in_channels', kw['in_channels']))][:nin]
        xs = [torch.randn(2, c, 9, 8, 7, requires_grad=True) for c in chans]
        y = op(xs[0] if nin == 1 else xs)
        g = torch.randn(y.shape)
        (y * g).sum().backward()
        yo = orc.hartley_mha(xs[0].detach(), op.weight_query.detach(), op.weight_key.detach(), op.weight_value.detach(),
                             op.weight_out.detach(), kw['num_modes'], kw.get('patch_size'),
                             key=xs[1].detach() if nin > 1 else None, value=xs[2].detach() if nin > 2 else None)
        check(f'HartleyMultiHeadAttention {tag}', yo, y.detach())
        out.update({f'{tag}/y': y.detach().numpy(), f'{tag}/g': g.numpy(), f'{tag}/modes': np.array(kw['num_modes']),
                    f'{tag}/patch': np.array(kw.get('patch_size') or (0, 0, 0)),
                    f'{tag}/wq': op.weight_query.detach().numpy(), f'{tag}/wk': op.weight_key.detach().numpy(),
                    f'{tag}/wv': op.weight_value.detach().numpy(), f'{tag}/wo': op.weight_out.detach().numpy(),
                    f'{tag}/dwq': op.weight_query.grad.numpy(), f'{tag}/dwk': op.weight_key.grad.numpy(),
                    f'{tag}/dwv': op.weight_value.grad.numpy(), f'{tag}/dwo': op.weight_out.grad.numpy()})
        for i, x in enumerate(xs):
            out[f'{tag}/x{i}'] = x.detach().numpy()
            out[f'{tag}/dx{i}'] = x.grad.numpy()
    save('hartley_mha', **out)


def case_block():
    torch.manual_seed(13)
    out = {}
    for tag, cin in (('plain', 8), ('mapped', 16)):
        blk = ref_xs.HNOXSBlock(2, cin, 8, (2, 3, 3))
        blk.apply(ref.hnosegxs.init_weights_for_snn)
        x = torch.randn(2, cin, 9, 8, 7)
        y = blk(x)
        sd = {('layers.0.' + k): v for k, v in blk.state_dict().items()}
        check(f'HNOXSBlock {tag}', orc.xs_block(x, sd, 'layers.0.', 2, (2, 3, 3)), y.detach())
        out.update({f'x_{tag}': x.numpy(), f'y_{tag}': y.detach().numpy()})
        out.update({f'sd_{tag}/{k}': v.numpy() for k, v in sd.items()})
    save('block', **out)


def small_model(weights_type='shared'):
    torch.manual_seed(14)
    cfg = dict(in_channels=2, out_channels=3, filters=8, num_transform_blocks=[1, 2, 1, 2, 1, 2], num_modes=(2, 3, 3),
               weights_type=weights_type)
    model = ref.HNOSegXS(**cfg)
    return cfg, model


def case_model():
    out = {}
    for wt in ('shared', 'individual'):
        cfg, model = small_model(wt)
        torch.manual_seed(15)
        x = torch.randn(2, 2, 18, 16, 13)
        labels = torch.randint(0, 3, (2, 1, 18, 16, 13))
        logits = {}
        hook = model.conv_out.register_forward_hook(lambda m, i, o: logits.__setitem__('v', o.detach()))
        probs = model(x)
        hook.remove()
        sd = {k: v.detach() for k, v in model.state_dict().items()}
        o_probs, o_logits = orc.hnosegxs_forward(sd, x, cfg['num_transform_blocks'], cfg['num_modes'],
                                                 return_logits=True)
        check(f'HNOSegXS {wt} probs', o_probs, probs.detach())
        check(f'HNOSegXS {wt} logits', o_logits, logits['v'])
        out.update({f'{wt}/x': x.numpy(), f'{wt}/labels': labels.numpy().astype(np.uint8),
                    f'{wt}/probs': probs.detach().numpy(), f'{wt}/logits': logits['v'].numpy()})
        out.update({f'{wt}/sd/{k}': v.numpy() for k, v in sd.items()})
        onehot = torch.zeros(2, 3, 18, 16, 13).scatter_(1, labels, 1.0)
        check('to_categorical', orc.to_categorical(labels, 3), onehot, 1e-7)
        for lname in ('DiceLoss', 'PCCLoss'):
            model.zero_grad()
            loss = getattr(ref_losses, lname)()(model(x), onehot)
            loss.backward()
            grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
            o_loss, o_grads = orc.train_step(sd, x, labels, cfg['num_transform_blocks'], cfg['num_modes'], lname)
            check(f'{wt} {lname} value', o_loss, loss.detach(), 1e-6)
            for k in grads:
                check(f'{wt} {lname} grad {k}', o_grads[k], grads[k], 2e-4)
            out[f'{wt}/{lname}/loss'] = loss.detach().numpy()
            out.update({f'{wt}/{lname}/grad/{k}': v.numpy() for k, v in grads.items()})
    save('model_small', **out)


def case_hartley_mha_seg():
    """HartleyMHASeg (BASELINE config 5's architecture, architectures.py:432-508) in small, deep supervision on (its
    default): forward + Dice gradients."""
    torch.manual_seed(61)
    cfg = dict(in_channels=2, out_channels=3, filters=8, num_transform_blocks=2, num_heads=2, num_modes=(2, 4, 2),
               patch_size=(2, 2, 2))
    model = ref.HartleyMHASeg(**cfg)
    torch.manual_seed(62)
    x = torch.randn(2, 2, 18, 16, 13)
    labels = torch.randint(0, 3, (2, 1, 18, 16, 13))
    probs = model(x)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    o_probs = orc.hnoseg_forward(sd, x, 2, cfg['num_modes'], patch=cfg['patch_size'])
    check('HartleyMHASeg probs', o_probs, probs.detach())
    onehot = torch.zeros(2, 3, 18, 16, 13).scatter_(1, labels, 1.0)
    model.zero_grad()
    loss = ref_losses.DiceLoss()(model(x), onehot)
    loss.backward()
    o_loss, o_grads = orc.hnoseg_train_step(sd, x, labels, 2, cfg['num_modes'], 'DiceLoss', patch=cfg['patch_size'])
    check('HartleyMHASeg DiceLoss value', o_loss, loss.detach(), 1e-6)
    out = {'x': x.numpy(), 'labels': labels.numpy().astype(np.uint8), 'probs': probs.detach().numpy(),
           'DiceLoss/loss': loss.detach().numpy()}
    for k, p in model.named_parameters():
        check(f'HartleyMHASeg DiceLoss grad {k}', o_grads[k], p.grad, 2e-4)
        out[f'DiceLoss/grad/{k}'] = p.grad.numpy().copy()
    out.update({f'sd/{k}': v.numpy() for k, v in sd.items()})
    save('hartley_mha_seg_small', **out)


def case_fourier_operator():
    """FourierOperator with transform, shared weights (BASELINE config 3's layer): forward + gradients, incl. the
    clamp path (modes larger than half the grid) and an even grid."""
    from nets.fourier_operator import FourierOperator
    out = {}
    for tag, shape, modes, wt in (('a', (9, 8, 7), (2, 3, 3), 'shared'), ('b', (6, 7, 8), (5, 2, 9), 'shared'),
                                  ('c', (9, 8, 7), (2, 4, 3), 'individual'), ('d', (8, 13, 6), (4, 2, 3), 'individual')):
        torch.manual_seed(31)
        # 8 channels: what the CUDA pointwise kernels are instantiated for.  'c' / 'd': per-mode complex weights
        # (config_fno.ini), incl. axes that the two corners fill exactly (n == 2m)
        op = FourierOperator(8, 8, modes, weights_type=wt)
        x = torch.randn(2, 8, *shape, requires_grad=True)
        y = op(x)
        w = torch.randn(y.shape)
        (y * w).sum().backward()
        yo = orc.fourier_operator_with_transform(x.detach(), op.weight_real.detach(), op.weight_imag.detach(), modes)
        check(f'FourierOperator {tag}', yo, y.detach())
        out.update({f'{tag}/x': x.detach().numpy(), f'{tag}/w': w.numpy(), f'{tag}/y': y.detach().numpy(),
                    f'{tag}/wr': op.weight_real.detach().numpy(), f'{tag}/wi': op.weight_imag.detach().numpy(),
                    f'{tag}/dx': x.grad.numpy(), f'{tag}/dwr': op.weight_real.grad.numpy(),
                    f'{tag}/dwi': op.weight_imag.grad.numpy(), f'{tag}/modes': np.array(modes)})
    save('fourier_operator', **out)


def case_hnoseg(transform_type='Hartley', name='hnoseg_small', **extra):
    """NeuralOperatorSeg(transform_type='Hartley') = HNOSeg / ('Fourier') = FNOSeg: small model,
    forward + Dice gradients."""
    torch.manual_seed(21)
    cfg = dict(in_channels=2, out_channels=3, filters=8, num_transform_blocks=3, num_modes=(2, 3, 3),
               transform_type=transform_type, **extra)
    skip = extra.get('use_block_skip', True)
    model = ref.NeuralOperatorSeg(**cfg)
    torch.manual_seed(22)
    x = torch.randn(2, 2, 18, 16, 13)
    labels = torch.randint(0, 3, (2, 1, 18, 16, 13))
    logits = {}
    hook = model.conv_out.register_forward_hook(lambda m, i, o: logits.__setitem__('v', o.detach()))
    probs = model(x)
    hook.remove()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    o_probs, o_logits = orc.hnoseg_forward(sd, x, cfg['num_transform_blocks'], cfg['num_modes'], return_logits=True,
                                           use_block_skip=skip)
    check('HNOSeg probs', o_probs, probs.detach())
    check('HNOSeg logits', o_logits, logits['v'])
    out = {'x': x.numpy(), 'labels': labels.numpy().astype(np.uint8), 'probs': probs.detach().numpy(),
           'logits': logits['v'].numpy()}
    out.update({f'sd/{k}': v.numpy() for k, v in sd.items()})
    onehot = torch.zeros(2, 3, 18, 16, 13).scatter_(1, labels, 1.0)
    model.zero_grad()
    loss = ref_losses.DiceLoss()(model(x), onehot)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    o_loss, o_grads = orc.hnoseg_train_step(sd, x, labels, cfg['num_transform_blocks'], cfg['num_modes'], 'DiceLoss',
                                            use_block_skip=skip)
    check('HNOSeg DiceLoss value', o_loss, loss.detach(), 1e-6)
    for k in grads:
        check(f'HNOSeg DiceLoss grad {k}', o_grads[k], grads[k], 2e-4)
    out['DiceLoss/loss'] = loss.detach().numpy()
    out.update({f'DiceLoss/grad/{k}': v.numpy() for k, v in grads.items()})
    save(name, **out)


def case_losses():
    torch.manual_seed(16)
    p = torch.softmax(torch.randn(2, 4, 6, 5, 7), dim=1).requires_grad_(True)
    t = orc.to_categorical(torch.randint(0, 4, (2, 1, 6, 5, 7)), 4)
    out = {'p': p.detach().numpy(), 't': t.numpy()}
    for lname in ('DiceLoss', 'PCCLoss', 'ExpDiceLoss', 'CrossEntropyLoss'):
        # experiments/run.py:105-110: custom_losses first, torch.nn otherwise
        loss_fn = getattr(ref_losses, lname)() if hasattr(ref_losses, lname) else getattr(torch.nn, lname)()
        loss = loss_fn(p, t)
        (g,) = torch.autograd.grad(loss, p)
        check(f'{lname}', orc.LOSSES[lname](p.detach(), t), loss.detach(), 1e-6)
        out[f'{lname}/loss'] = loss.detach().numpy()
        out[f'{lname}/grad'] = g.numpy()
    save('losses', **out)


def _reference_experiments_utils():
    """experiments/utils.py needs SimpleITK and torchinfo at import time (absent here); only its two array
    helpers matter, so the module is loaded by path with those two imports stubbed."""
    import importlib.util
    import types
    for name in ('SimpleITK', 'torchinfo'):
        if name not in sys.modules:
            stub = types.ModuleType(name)
            stub.summary = None
            sys.modules[name] = stub
    spec = importlib.util.spec_from_file_location('_ref_experiments_utils', os.path.join('/root/reference', 'experiments', 'utils.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def case_input_side():
    """to_categorical / normalize_modalities recorded from the reference's experiments/utils.py."""
    ru = _reference_experiments_utils()
    rng = np.random.default_rng(41)
    out = {}
    labels = torch.from_numpy(rng.integers(0, 4, (2, 1, 6, 5, 7)))
    onehot = ru.to_categorical(labels, 4)
    assert torch.equal(orc.to_categorical(labels, 4), onehot)
    out['labels'] = labels.numpy().astype(np.uint8)
    out['onehot'] = np.ascontiguousarray(onehot.numpy())
    # BraTS-like intensities: background exactly 0, tissue in the hundreds / thousands, four modalities of different scale
    vol = rng.gamma(4.0, 150.0, (4, 9, 10, 11)).astype(np.float32) * np.array([1, 0.5, 3, 0.1], np.float32).reshape(4, 1, 1, 1)
    vol[:, :2] = 0
    vol[:, :, :, -3:] = 0
    out['vol'] = vol
    for tag, kw in (('plain', {}), ('mask', dict(mask_val=0)), ('clip', dict(clip_val=(50.0, 900.0))),
                    ('maskclip', dict(mask_val=0, clip_val=(0.0, 700.0))), ('maskhit', dict(mask_val=700, clip_val=(0.0, 700.0)))):
        y = ru.normalize_modalities(vol, **kw)
        check(f'normalize_modalities {tag}', torch.from_numpy(orc.normalize_modalities(vol, **kw)), torch.from_numpy(y), 1e-5)
        out[f'norm/{tag}'] = y
    save('input_side', **out)


def _reference_image_transform():
    """experiments/data_io/dataset.py imports SimpleITK (absent here).  The module is loaded by path with a stand-in
    `SimpleITK` whose AffineTransform records (matrix, offset) and whose ResampleImageFilter.Execute runs the oracle's
    nearest-neighbour resampler: everything the REFERENCE computes (parameter draws, matrix composition, centring, channel
    loop, flips) runs unmodified; only the third-party resampler is the restatement (oracle header: parity unpinned there)."""
    import importlib.util
    import types
    captured = []

    class AffineTransform:
        def __init__(self, matrix, offset):
            self.matrix = np.asarray(matrix, dtype=np.float64)
            self.offset = np.asarray(offset, dtype=np.float64)
            captured.append((self.matrix.copy(), self.offset.copy()))

    class _Image:
        def __init__(self, arr):
            self.arr = np.asarray(arr)

        def GetSize(self):
            return self.arr.shape[::-1]

        def GetSpacing(self):
            return (1.0,) * self.arr.ndim

        def GetOrigin(self):
            return (0.0,) * self.arr.ndim

    class ResampleImageFilter:
        def SetInterpolator(self, interp):
            assert interp == 'nn'

        def SetDefaultPixelValue(self, v):
            self.cval = v

        def SetTransform(self, t):
            self.t = t

        def SetSize(self, s):
            self.size = s

        def SetOutputSpacing(self, s):
            assert all(v == 1.0 for v in s)

        def SetOutputOrigin(self, o):
            assert all(v == 0.0 for v in o)

        def Execute(self, image):
            assert tuple(self.size) == tuple(image.GetSize())
            n = image.arr.ndim
            out = orc.affine_resample_nn(image.arr[None], self.t.matrix.reshape(n, n), self.t.offset, self.cval)[0]
            return _Image(out)

    stub = types.ModuleType('SimpleITK')
    stub.AffineTransform = AffineTransform
    stub.ResampleImageFilter = ResampleImageFilter
    stub.sitkNearestNeighbor = 'nn'
    stub.GetImageFromArray = lambda a: _Image(a)
    stub.GetArrayFromImage = lambda im: im.arr
    saved = sys.modules.get('SimpleITK')
    sys.modules['SimpleITK'] = stub
    try:
        spec = importlib.util.spec_from_file_location('_ref_dataset', os.path.join('/root/reference', 'experiments', 'data_io', 'dataset.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules['SimpleITK'] = saved
    return mod, captured


AUGMENT_CASES = orc.AUGMENT_CASES


def case_augment():
    """ImageTransform (experiments/data_io/dataset.py:63-192) recorded from the reference: for every case the inputs, the
    (matrix, offset) it hands to sitk.AffineTransform per call and the augmented image / label arrays."""
    ref_ds, captured = _reference_image_transform()
    out = {}
    for name, spatial, kw, calls in AUGMENT_CASES:
        data_rng = np.random.default_rng(100 + len(name))
        x = data_rng.normal(size=(3,) + spatial).astype(np.float32)
        y = data_rng.integers(0, 4, (1,) + spatial).astype(np.uint8)
        tr = ref_ds.ImageTransform(**kw)
        okw = {k: v for k, v in kw.items() if k != 'seed'}
        orng = np.random.default_rng(kw.get('seed'))
        out[f'{name}/x'], out[f'{name}/y'] = x, y
        for c in range(calls):
            del captured[:]
            xr, yr = tr(x, y)
            xo, yo, rec = orc.image_transform(x, y, orng, **okw)
            assert np.array_equal(np.asarray(xr), np.asarray(xo)) and np.array_equal(np.asarray(yr), np.asarray(yo)), (name, c)
            if captured:
                # one AffineTransform per apply_transform call (x and y): identical
                m, o = captured[0]
                nd = len(spatial)
                assert np.allclose(m.reshape(nd, nd), rec['matrix'], rtol=0, atol=1e-12) and np.allclose(o, rec['offset'], rtol=0, atol=1e-9)
                out[f'{name}/{c}/matrix'], out[f'{name}/{c}/offset'] = m.reshape(nd, nd), o
            else:
                assert rec['matrix'] is None
            out[f'{name}/{c}/flips'] = np.asarray(rec['flips'])
            out[f'{name}/{c}/xo'] = np.ascontiguousarray(xr)
            out[f'{name}/{c}/yo'] = np.ascontiguousarray(yr)
        print(f'  augment {name}: {calls} calls identical (reference ImageTransform vs oracle)')
    save('augment', **out)


def case_noresize():
    """use_resize=False (nets/hnosegxs.py:102-109, 150, 174; nets/architectures.py:286-289, 345-347): the blocks run at the
    image resolution, no stem and no interpolation.  HNOSegXS with 2 input channels on a grid whose planes need padding and
    with 4 channels on one that does not, and NeuralOperatorSeg('Hartley') with deep supervision: forward + Dice gradients."""
    out = {}
    xs_cases = (('xs2', 2, (11, 10, 9), 31), ('xs4', 4, (10, 8, 8), 32))
    for tag, cin, spatial, seed in xs_cases:
        torch.manual_seed(seed)
        cfg = dict(in_channels=cin, out_channels=3, filters=8, num_transform_blocks=[1, 2, 1, 2], num_modes=(2, 3, 3),
                   use_resize=False)
        model = ref.HNOSegXS(**cfg)
        x = torch.randn(2, cin, *spatial)
        labels = torch.randint(0, 3, (2, 1) + spatial)
        onehot = torch.zeros(2, 3, *spatial).scatter_(1, labels, 1.0)
        sd = {k: v.detach() for k, v in model.state_dict().items()}
        assert 'conv_in.op.weight' not in sd
        probs = model(x)
        loss = ref_losses.DiceLoss()(probs, onehot)
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        check(f'{tag} probs', orc.hnosegxs_forward(sd, x, cfg['num_transform_blocks'], cfg['num_modes'], use_resize=False),
              probs.detach())
        o_loss, o_grads = orc.train_step(sd, x, labels, cfg['num_transform_blocks'], cfg['num_modes'], 'DiceLoss')
        check(f'{tag} DiceLoss value', o_loss, loss.detach(), 1e-6)
        for k in grads:
            check(f'{tag} grad {k}', o_grads[k], grads[k], 2e-4)
        out.update({f'{tag}/x': x.numpy(), f'{tag}/labels': labels.numpy().astype(np.uint8), f'{tag}/probs': probs.detach().numpy(),
                    f'{tag}/loss': loss.detach().numpy()})
        out.update({f'{tag}/sd/{k}': v.numpy() for k, v in sd.items()})
        out.update({f'{tag}/grad/{k}': v.numpy() for k, v in grads.items()})
    torch.manual_seed(33)
    cfg = dict(in_channels=3, out_channels=3, filters=8, num_transform_blocks=2, num_modes=(2, 3, 3), transform_type='Hartley',
               use_resize=False, use_deep_supervision=True)
    model = ref.NeuralOperatorSeg(**cfg)
    spatial = (10, 9, 12)
    x = torch.randn(2, 3, *spatial)
    labels = torch.randint(0, 3, (2, 1) + spatial)
    onehot = torch.zeros(2, 3, *spatial).scatter_(1, labels, 1.0)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    probs = model(x)
    loss = ref_losses.DiceLoss()(probs, onehot)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    check('hnoseg probs', orc.hnoseg_forward(sd, x, 2, cfg['num_modes']), probs.detach())
    o_loss, o_grads = orc.hnoseg_train_step(sd, x, labels, 2, cfg['num_modes'], 'DiceLoss')
    check('hnoseg DiceLoss value', o_loss, loss.detach(), 1e-6)
    for k in grads:
        check(f'hnoseg grad {k}', o_grads[k], grads[k], 2e-4)
    out.update({'hnoseg/x': x.numpy(), 'hnoseg/labels': labels.numpy().astype(np.uint8), 'hnoseg/probs': probs.detach().numpy(),
                'hnoseg/loss': loss.detach().numpy()})
    out.update({f'hnoseg/sd/{k}': v.numpy() for k, v in sd.items()})
    out.update({f'hnoseg/grad/{k}': v.numpy() for k, v in grads.items()})
    save('noresize_small', **out)


def case_full():
    """BASELINE config 1: HNOSegXS(4,4,24,[3]*8,(10,14,14)) on one 1x4x240x240x155 volume."""
    torch.manual_seed(0)
    model = ref.HNOSegXS(4, 4, 24, [3] * 8, (10, 14, 14))
    n_params = sum(p.numel() for p in model.parameters())
    assert n_params == 28248, n_params  # README.md:57-63
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, 4, 240, 240, 155, generator=g)
    logits = {}
    model.conv_out.register_forward_hook(lambda m, i, o: logits.__setitem__('v', o.detach()))
    with torch.no_grad():
        probs = model(x)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    with torch.no_grad():
        o_probs, o_logits = orc.hnosegxs_forward(sd, x, [3] * 8, (10, 14, 14), return_logits=True)
    lg = logits['v']
    rel = ((o_logits - lg).norm() / lg.norm()).item()
    agree = (o_logits.argmax(1) == lg.argmax(1)).float().mean().item()
    print(f'  full size: oracle vs reference logits rel-L2 {rel:.2e}, argmax agreement {agree:.7f}')
    assert rel < 1e-5 and agree > 0.99999
    gi = torch.Generator().manual_seed(77)
    idx = torch.randint(0, 240 * 240 * 155, (8192,), generator=gi)
    flat = lg.reshape(4, -1)
    out = {'idx': idx.numpy(), 'logits_at_idx': flat[:, idx].numpy(), 'probs_at_idx': probs.reshape(4, -1)[:, idx].numpy(),
           'logits_std': np.float32(lg.std().item()), 'logits_norm': np.float32(lg.norm().item()),
           'argmax_hist': np.bincount(lg.argmax(1).flatten().numpy(), minlength=4)}
    out.update({f'sd/{k}': v.numpy() for k, v in sd.items()})
    save('model_full_probe', **out)


def case_deep_supervision():
    """use_deep_supervision=True (nets/hnosegxs.py:110-125, 154-172; nets/architectures.py:295-311, 330-343) for HNOSeg-XS
    and HNOSeg in small: probabilities, Dice loss and every parameter gradient recorded from the real reference (the
    HartleyMHASeg fixture covers the third user of the same head)."""
    out = {}
    torch.manual_seed(31)
    xs = ref.HNOSegXS(2, 3, 8, [1, 2, 1, 2, 1, 2], (2, 3, 3), use_deep_supervision=True)
    torch.manual_seed(32)
    hn = ref.NeuralOperatorSeg(2, 3, 8, 3, (2, 3, 3), 'Hartley', use_deep_supervision=True)
    torch.manual_seed(33)
    x = torch.randn(2, 2, 18, 16, 13)
    labels = torch.randint(0, 3, (2, 1, 18, 16, 13))
    onehot = torch.zeros(2, 3, 18, 16, 13).scatter_(1, labels, 1.0)
    out.update({'x': x.numpy(), 'labels': labels.numpy().astype(np.uint8)})
    for tag, model in (('xs', xs), ('hnoseg', hn)):
        model.zero_grad()
        probs = model(x)
        loss = ref_losses.DiceLoss()(probs, onehot)
        loss.backward()
        out[f'{tag}/probs'] = probs.detach().numpy()
        out[f'{tag}/DiceLoss/loss'] = loss.detach().numpy()
        out.update({f'{tag}/sd/{k}': v.detach().numpy().copy() for k, v in model.state_dict().items()})
        out.update({f'{tag}/DiceLoss/grad/{k}': p.grad.numpy().copy() for k, p in model.named_parameters()})
        print(f'  deep supervision {tag}: loss {float(loss):.6f}, conv_out {tuple(model.conv_out.weight.shape)}')
    # the oracle restates the _TransSeg head (used for HartleyMHASeg): pin it on the HNOSeg variant too
    sd = {k: v.detach() for k, v in hn.state_dict().items()}
    check('HNOSeg deep supervision probs', orc.hnoseg_forward(sd, x, 3, (2, 3, 3)), torch.from_numpy(out['hnoseg/probs']))
    save('deep_supervision_small', **out)


def case_superres():
    """BASELINE config 4: zero-shot super-resolution = the SAME weights on a 2x grid (README.md:83-87), run through the
    reference's testing path (experiments/train_test.py:373-414: model.eval(), no_grad, forward, host argmax).  Stores the
    reference's logits / probabilities / label map at 8192 sampled voxels of the 1x4x480x480x310 volume."""
    torch.manual_seed(0)
    model = ref.HNOSegXS(4, 4, 24, [3] * 8, (10, 14, 14))
    model.eval()
    x = torch.randn(1, 4, 480, 480, 310, generator=torch.Generator().manual_seed(9))
    logits = {}
    model.conv_out.register_forward_hook(lambda m, i, o: logits.__setitem__('v', o.detach()))
    with torch.no_grad():
        probs = model(x)
    lg = logits['v']
    labels = np.argmax(probs.numpy(), axis=1).astype(np.uint8)  # train_test.py:402-408
    idx = torch.randint(0, 480 * 480 * 310, (8192,), generator=torch.Generator().manual_seed(78))
    out = {'idx': idx.numpy(), 'logits_at_idx': lg.reshape(4, -1)[:, idx].numpy(),
           'probs_at_idx': probs.reshape(4, -1)[:, idx].numpy(), 'labels_at_idx': labels.reshape(-1)[idx.numpy()],
           'logits_norm': np.float32(lg.norm().item()), 'label_hist': np.bincount(labels.reshape(-1), minlength=4)}
    # top-2 logit margin of every sampled voxel (the GPU test may only disagree where the margin is within round-off)
    top2 = lg.reshape(4, -1)[:, idx].topk(2, dim=0).values
    out['margin_at_idx'] = (top2[0] - top2[1]).numpy()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    full = dict(np.load(os.path.join(GOLDEN, 'model_full_probe.npz')))
    for k, v in sd.items():  # same weights as the 1x probe: the fixture does not repeat them
        assert np.array_equal(full[f'sd/{k}'], v.numpy()), k
    del probs, lg
    with torch.no_grad():
        _, o_logits = orc.hnosegxs_forward(sd, x, [3] * 8, (10, 14, 14), return_logits=True)
    o_at = o_logits.reshape(4, -1)[:, idx]
    rel = ((o_at - torch.from_numpy(out['logits_at_idx'])).norm() / torch.from_numpy(out['logits_at_idx']).norm()).item()
    print(f'  2x grid: oracle vs reference logits rel-L2 at the probe {rel:.2e}')
    assert rel < 1e-5
    save('model_superres_probe', **out)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--full', action='store_true')
    ap.add_argument('--only', default=None, help='run a single case function by name, e.g. case_superres')
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(8)
    if args.only:
        globals()[args.only]()
        sys.exit(0)
    case_dht()
    case_operator()
    case_operator_transform_individual()
    case_hartley_mha()
    case_hartley_mha_seg()
    case_block()
    case_losses()
    case_input_side()
    case_augment()
    case_noresize()
    case_model()
    case_hnoseg()
    case_fourier_operator()
    case_hnoseg('Fourier', 'fnoseg_small')
    case_hnoseg('Hartley', 'hnoseg_individual_small', weights_type='individual')
    # config_fno.ini: the original FNO (per-mode complex weights, biased conv branch, no block skip)
    case_hnoseg('Fourier', 'fno_small', weights_type='individual', use_bias_conv_branch=True, use_block_skip=False)
    case_deep_supervision()
    if args.full:
        case_full()
        case_superres()
    print('all oracle-vs-reference checks passed')
