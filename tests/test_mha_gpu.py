"""Hartley multi-head attention, HartleyMHASeg and deep supervision on the CUDA path (BASELINE config 5)
against fixtures recorded from the real reference (oracle/make_golden.py) and against the live oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import hno_oracle as orc

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


@pytest.fixture
def tensor_cores():
    from multimodal_3d_image_segmentation_b200 import _lib
    lib = _lib.load()
    yield lib.hno_set_tensor_cores
    lib.hno_set_tensor_cores(1)


CASES = {'self_grouped': (dict(in_channels=8, key_dim=6, num_heads=2, num_modes=(2, 4, 2), patch_size=(2, 2, 2)), 1),
         'self_plain': (dict(in_channels=8, key_dim=5, num_heads=3, num_modes=(2, 3, 3), value_dim=4), 1),
         'cross': (dict(in_channels=8, key_dim=4, num_heads=2, num_modes=(2, 2, 3), patch_size=(1, 2, 3),
                        key_in_channels=6, value_in_channels=5), 3)}


@pytest.mark.parametrize('tc', [1, 0])
@pytest.mark.parametrize('tag', list(CASES))
def test_hartley_mha_layer_against_reference_fixture(cuda, golden_dir, tensor_cores, tag, tc):
    """HartleyMultiHeadAttention (hartley_mha.py:136-222): grouped / plain self-attention, value_dim != key_dim and
    cross-attention with separate key / value inputs; output and every gradient recorded from the real reference.  Both the
    tcgen05 GEMM route and the CUDA-core route."""
    from multimodal_3d_image_segmentation_b200 import nets
    tensor_cores(tc)
    g = dict(np.load(os.path.join(golden_dir, 'hartley_mha.npz')))
    kw, nin = CASES[tag]
    op = nets.HartleyMultiHeadAttention(**kw, device=cuda)
    with torch.no_grad():
        op.weight_query.copy_(torch.from_numpy(g[f'{tag}/wq']))
        op.weight_key.copy_(torch.from_numpy(g[f'{tag}/wk']))
        op.weight_value.copy_(torch.from_numpy(g[f'{tag}/wv']))
        op.weight_out.copy_(torch.from_numpy(g[f'{tag}/wo']))
    xs = [torch.from_numpy(g[f'{tag}/x{i}']).to(cuda).requires_grad_(True) for i in range(nin)]
    y = op(xs[0] if nin == 1 else xs)
    assert rel(y, g[f'{tag}/y']) < 1e-5, (tag, rel(y, g[f'{tag}/y']))
    (y * torch.from_numpy(g[f'{tag}/g']).to(cuda)).sum().backward()
    for i, x in enumerate(xs):
        assert rel(x.grad, g[f'{tag}/dx{i}']) < 2e-5, (tag, i, rel(x.grad, g[f'{tag}/dx{i}']))
    for name, p in (('dwq', op.weight_query), ('dwk', op.weight_key), ('dwv', op.weight_value), ('dwo', op.weight_out)):
        assert rel(p.grad, g[f'{tag}/{name}']) < 2e-5, (tag, name, rel(p.grad, g[f'{tag}/{name}']))


def test_hartley_mha_bias_and_no_activation_against_oracle(cuda):
    """use_bias=True (biases on Q / K / V / output, hartley_mha.py:174-177, 217-218) and attention_activation=None, on
    already-transformed inputs (use_transform=False, :224-296), against a plain-torch restatement in fp64."""
    from multimodal_3d_image_segmentation_b200 import nets
    gen = torch.Generator().manual_seed(3)
    for act in ('selu', None):
        op = nets.HartleyMultiHeadAttention(6, 4, 2, (2, 3, 2), (2, 2, 1), attention_activation=act, use_bias=True,
                                            use_transform=False, device=cuda)
        with torch.no_grad():
            for p in op.parameters():
                p.copy_(torch.randn(p.shape, generator=gen) * 0.4)
        z = torch.randn(2, 6, 4, 6, 4, generator=gen)
        zc = z.to(cuda).requires_grad_(True)
        y = op(zc)
        w = torch.randn(y.shape, generator=gen)
        (y * w.to(cuda)).sum().backward()
        # restatement (fp64)
        P = {k: v.detach().cpu().double().requires_grad_(True) for k, v in op.named_parameters()}
        zr = z.double().requires_grad_(True)
        q = torch.einsum('zoi,bidhw->bzodhw', P['weight_query'], zr) + P['bias_query']
        k = torch.einsum('zoi,bidhw->bzodhw', P['weight_key'], zr) + P['bias_key']
        v = torch.einsum('zoi,bidhw->bzodhw', P['weight_value'], zr) + P['bias_value']
        q, k, v = (orc.group_patches(t, (2, 2, 1)) for t in (q, k, v))
        grid = q.shape[3:]
        q, k, v = (t.flatten(3) for t in (q, k, v))
        att = torch.einsum('bzcq,bzck->bzqk', q, k) / np.sqrt(k.shape[2])
        if act is not None:
            att = torch.nn.functional.selu(att)
        o = torch.einsum('bzqk,bzck->bzcq', att, v).reshape(v.shape[:3] + tuple(grid))
        o = orc.ungroup_patches(o, 4, (2, 2, 1)).flatten(1, 2)
        yr = torch.einsum('oi,bidhw->bodhw', P['weight_out'], o) + P['bias_out']
        (yr * w.double()).sum().backward()
        assert rel(y, yr) < 1e-5, (act, rel(y, yr))
        assert rel(zc.grad, zr.grad) < 2e-5
        for name, p in op.named_parameters():
            assert rel(p.grad, P[name].grad) < 2e-5, (act, name, rel(p.grad, P[name].grad))


@pytest.mark.parametrize('tc', [1, 0])
def test_hartley_mha_seg_against_reference_fixture(cuda, golden_dir, tensor_cores, tc):
    """HartleyMHASeg in small with its default deep supervision (architectures.py:432-508, 295-311): probabilities, Dice
    loss and every parameter gradient of the real reference, through the fused engine."""
    from multimodal_3d_image_segmentation_b200 import nets
    from multimodal_3d_image_segmentation_b200.engine import TransSegEngine
    tensor_cores(tc)
    g = dict(np.load(os.path.join(golden_dir, 'hartley_mha_seg_small.npz')))
    model = nets.HartleyMHASeg(2, 3, 8, 2, 2, (2, 4, 2), (2, 2, 2), device=cuda)
    sd = _sd(g, 'sd/')
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(sd)
    assert TransSegEngine.supports(model)
    x = torch.from_numpy(g['x']).to(cuda)
    labels = torch.from_numpy(g['labels'].astype(np.int64))
    probs = model(x)
    assert rel(probs, g['probs']) < 1e-5, rel(probs, g['probs'])
    loss = nets.custom_losses.DiceLoss()(probs, orc.to_categorical(labels, 3).to(cuda))
    loss.backward()
    assert abs(float(loss.detach()) - float(g['DiceLoss/loss'])) < 2e-6
    for k, p in model.named_parameters():
        ref = g[f'DiceLoss/grad/{k}']
        assert rel(p.grad, ref) < 2e-4, (k, rel(p.grad, ref))


def test_deep_supervision_against_reference_fixture(cuda, golden_dir):
    """use_deep_supervision=True for HNOSeg-XS (nets/hnosegxs.py:110-125, 154-172: conv_out over all nine outputs) and HNOSeg
    (architectures.py:306-311, 339-343: conv_ds + SELU, then conv_out): probabilities, Dice loss and all gradients of the real
    reference; also through the fused loss on integer labels and the Trainer's flat gradient."""
    from multimodal_3d_image_segmentation_b200 import nets, parallel
    g = dict(np.load(os.path.join(golden_dir, 'deep_supervision_small.npz')))
    x = torch.from_numpy(g['x']).to(cuda)
    labels = torch.from_numpy(g['labels'].astype(np.int64))
    onehot = orc.to_categorical(labels, 3).to(cuda)
    models = {'xs': nets.HNOSegXS(2, 3, 8, [1, 2, 1, 2, 1, 2], (2, 3, 3), use_deep_supervision=True, device=cuda),
              'hnoseg': nets.NeuralOperatorSeg(2, 3, 8, 3, (2, 3, 3), 'Hartley', use_deep_supervision=True, device=cuda)}
    for tag, model in models.items():
        sd = _sd(g, f'{tag}/sd/')
        assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v.shape) for k, v in model.state_dict().items()}
        model.load_state_dict(sd)
        probs = model(x)
        assert rel(probs, g[f'{tag}/probs']) < 1e-5, (tag, rel(probs, g[f'{tag}/probs']))
        loss = nets.custom_losses.DiceLoss()(probs, onehot)
        loss.backward()
        assert abs(float(loss.detach()) - float(g[f'{tag}/DiceLoss/loss'])) < 2e-6
        for k, p in model.named_parameters():
            ref = g[f'{tag}/DiceLoss/grad/{k}']
            assert rel(p.grad, ref) < 2e-4, (tag, k, rel(p.grad, ref))
    xs = models['xs']
    xs.zero_grad()
    loss = xs.loss(x, labels.to(cuda), 'DiceLoss')
    loss.backward()
    assert abs(float(loss.detach()) - float(g['xs/DiceLoss/loss'])) < 2e-6
    for k, p in xs.named_parameters():
        assert rel(p.grad, g[f'xs/DiceLoss/grad/{k}']) < 2e-4, k
    tr = parallel.Trainer(xs, 'DiceLoss', use_graph=False)
    tr.loss_and_grad(x, labels.to(cuda))
    for k, p in xs.named_parameters():
        assert rel(tr.flat.grad_view_of(p), g[f'xs/DiceLoss/grad/{k}']) < 2e-4, k


@pytest.mark.parametrize('tc', [1, 0])
def test_hartley_mha_layer_baseline_config_against_oracle(cuda, tensor_cores, tc):
    """BASELINE config 5's layer at its real size: 12 channels on the 121 x 121 x 78 grid, modes (10, 14, 14), patch 2 ->
    1,960 tokens x 96 features x 4 heads; forward and all gradients against the oracle."""
    from multimodal_3d_image_segmentation_b200 import nets
    tensor_cores(tc)
    gen = torch.Generator().manual_seed(11)
    op = nets.HartleyMultiHeadAttention(12, 12, 4, (10, 14, 14), (2, 2, 2), device=cuda)
    with torch.no_grad():
        for p in op.parameters():
            p.copy_(torch.randn(p.shape, generator=gen) * 0.3)
    x = torch.randn(1, 12, 121, 121, 78, generator=gen)
    xc = x.to(cuda).requires_grad_(True)
    y = op(xc)
    w = torch.randn(y.shape, generator=gen)
    (y * w.to(cuda)).sum().backward()
    P = {k: v.detach().cpu().double().requires_grad_(True) for k, v in op.named_parameters()}
    xr = x.double().requires_grad_(True)
    yr = orc.hartley_mha(xr, P['weight_query'], P['weight_key'], P['weight_value'], P['weight_out'], (10, 14, 14), (2, 2, 2))
    (yr * w.double()).sum().backward()
    # the fp32 reference's own distance from fp64 on this (un-normalised, 1,960-token) problem sets the scale
    P32 = {k: v.detach().cpu().requires_grad_(True) for k, v in op.named_parameters()}
    x32 = x.clone().requires_grad_(True)
    y32 = orc.hartley_mha(x32, P32['weight_query'], P32['weight_key'], P32['weight_value'], P32['weight_out'], (10, 14, 14),
                          (2, 2, 2))
    (y32 * w).sum().backward()
    worst = max((rel(p.grad, P[name].grad), name) for name, p in op.named_parameters())
    worst32 = max((rel(P32[name].grad, P[name].grad), name) for name in P)
    print(f'tc={tc}: y rel {rel(y, yr):.2e} (fp32 oracle {rel(y32, yr):.2e}), dx rel {rel(xc.grad, xr.grad):.2e} '
          f'(fp32 oracle {rel(x32.grad, xr.grad):.2e}), worst weight gradient {worst} (fp32 oracle {worst32})')
    assert rel(y, yr) < 5e-5
    assert rel(xc.grad, xr.grad) < 3e-4
    assert worst[0] < 3e-4, worst


def test_hartley_mha_seg_baseline_config_half_volume(cuda):
    """HartleyMHASeg(4, 4, 12, 16, 4, (10, 14, 14), (2, 2, 2)) (BASELINE config 5; hyper-parameters of the reference's
    config_hartleymha.ini) on a half-size volume: probabilities against the fp32 oracle, Dice gradients against the fp64 one."""
    from multimodal_3d_image_segmentation_b200 import nets
    torch.manual_seed(5)
    model = nets.HartleyMHASeg(4, 4, 12, 16, 4, (10, 14, 14), (2, 2, 2), device=cuda)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    x = torch.randn(1, 4, 120, 112, 77, generator=torch.Generator().manual_seed(6))
    labels = torch.randint(0, 4, (1, 1, 120, 112, 77), generator=torch.Generator().manual_seed(7))
    probs = model(x.to(cuda))
    loss = nets.custom_losses.DiceLoss()(probs, orc.to_categorical(labels, 4).to(cuda))
    loss.backward()
    with torch.no_grad():
        o_probs, o_logits = orc.hnoseg_forward(sd, x, 16, (10, 14, 14), return_logits=True, patch=(2, 2, 2))
    r = rel(probs, o_probs)
    agree = (probs.argmax(1).cpu() == o_probs.argmax(1)).float().mean().item()
    sd64 = {k: v.double() for k, v in sd.items()}
    o_loss, o_grads = orc.hnoseg_train_step(sd64, x.double(), labels, 16, (10, 14, 14), 'DiceLoss', patch=(2, 2, 2))
    flat = torch.cat([p.grad.flatten().cpu().double() for _, p in model.named_parameters()])
    oflat = torch.cat([o_grads[k].flatten() for k, _ in model.named_parameters()])
    rg = ((flat - oflat).norm() / oflat.norm()).item()
    print(f'HartleyMHASeg half volume: probs rel-L2 {r:.2e}, argmax agreement {agree:.6f}, flat gradient rel-L2 {rg:.2e}')
    assert r < 1e-3 and agree >= 0.9999
    assert abs(float(loss.detach()) - float(o_loss)) < 1e-5
    assert rg < 2e-3
