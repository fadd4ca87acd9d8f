"""Parity of every CUDA entry point (called through the C ABI) against the CPU oracle / fp64 torch maths."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import hno_oracle as orc

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def maxrel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def to_planar(x, pitch):
    """dense (B,C,D,H,W) -> planar (B,C,D,pitch) with zero padding columns"""
    B, C, D, H, W = x.shape
    out = torch.zeros(B, C, D, pitch, dtype=x.dtype, device=x.device)
    out[..., :H * W] = x.reshape(B, C, D, H * W)
    return out


def from_planar(x, H, W):
    B, C, D, P = x.shape
    return x[..., :H * W].reshape(B, C, D, H, W)


# ---------------------------------------------------------------------------------------------- DHT
DHT_CASES = [((9, 8, 7), (2, 3, 3)), ((12, 10, 9), (10, 14, 14)), ((16, 11, 10), (4, 3, 5)), ((5, 5, 5), (1, 2, 1)),
             ((40, 40, 26), (10, 14, 14)), ((31, 33, 20), (10, 14, 14))]


@pytest.mark.parametrize('shape,modes', DHT_CASES)
@pytest.mark.parametrize('padded', [False, True])
def test_dht_forward_adjoint(cuda, shape, modes, padded):
    from multimodal_3d_image_segmentation_b200 import ops
    from multimodal_3d_image_segmentation_b200.plan import get_crop_plan, plane_pitch
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, *shape, generator=g)
    plan = get_crop_plan(shape, modes, cuda)
    n = float(np.prod(shape))
    z_ref = orc.transform_crop(x.double(), modes)
    H, W = shape[1:]
    pitch = plane_pitch(H, W) if padded else None
    xd = to_planar(x.to(cuda), pitch) if padded else x.to(cuda)
    z = ops.dht3_forward(xd, plan, 1.0 / n)
    assert rel(z, z_ref) < 2e-6
    zz = torch.randn(z_ref.shape, generator=g)
    y_ref = orc.pad_inverse(zz.double(), shape)
    y = ops.dht3_adjoint(zz.to(cuda), plan, 1.0, pitch=pitch)
    if padded:
        assert float(y[..., H * W:].abs().max()) == 0.0 if pitch > H * W else True
        y = from_planar(y, H, W)
    assert rel(y, y_ref) < 2e-6
    # fused SELU epilogue and accumulate epilogue
    ys = ops.dht3_adjoint(zz.to(cuda), plan, 1.0, epilogue=2, pitch=pitch)
    ys = from_planar(ys, H, W) if padded else ys
    assert rel(ys, F.selu(y_ref)) < 2e-6
    base = torch.randn(2, 3, *shape, generator=g)
    acc = to_planar(base.to(cuda), pitch) if padded else base.to(cuda).clone()
    ops.dht3_adjoint(zz.to(cuda), plan, 0.5, epilogue=1, out=acc)
    acc = from_planar(acc, H, W) if padded else acc
    assert rel(acc, base.double() + 0.5 * y_ref) < 2e-6


def test_dht_golden_fixtures(cuda, golden_dir):
    """TransformCrop / PadInverse outputs recorded from the real reference (tests/golden/dht.npz)."""
    from multimodal_3d_image_segmentation_b200 import nets
    g = dict(np.load(os.path.join(golden_dir, 'dht.npz')))
    for tag in 'abc':
        x = torch.from_numpy(g[f'x_{tag}']).to(cuda)
        modes = tuple(int(v) for v in g[f'modes_{tag}'])
        z = nets.TransformCrop(modes, 5)(x)
        assert maxrel(z, g[f'z_{tag}']) < 1e-5
        y = nets.PadInverse(5)(torch.from_numpy(g[f'z_{tag}']).to(cuda), x.shape[2:])
        assert maxrel(y, g[f'y_{tag}']) < 1e-5
    x = torch.from_numpy(g['x']).to(cuda)
    assert maxrel(nets.dhtn(x, dim=(-3, -2, -1)), g['fwd']) < 1e-5
    assert maxrel(nets.dht3(x, is_inverse=True), g['inv']) < 1e-5


def test_dht_autograd(cuda):
    from multimodal_3d_image_segmentation_b200 import nets
    shape, modes = (9, 8, 7), (2, 3, 3)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(1, 2, *shape, generator=g)
    w = torch.randn(1, 2, 4, 6, 6, generator=g)
    xr = x.double().requires_grad_(True)
    zr = orc.transform_crop(xr, modes)
    yr = orc.pad_inverse(zr * zr, shape)
    (yr * yr).sum().backward()
    xc = x.to(cuda).requires_grad_(True)
    z = nets.TransformCrop(modes, 5)(xc)
    y = nets.PadInverse(5)(z * z, shape)
    (y * y).sum().backward()
    assert rel(y, yr) < 5e-6 and rel(xc.grad, xr.grad) < 5e-6
    del w


# ---------------------------------------------------------------------------------------------- pointwise conv
def _pw_ref(in1, in2, w, b, act, res):
    x = in1 if in2 is None else torch.cat([in1, in2], 1)
    y = torch.einsum('oi,bis->bos', w, x)
    if b is not None:
        y = y + b.view(1, -1, 1)
    if res:
        y = y + in1
    return F.selu(y) if act else y


@pytest.mark.parametrize('ci1,ci2,co,act,res,bias', [(24, 0, 24, 1, True, False), (24, 0, 24, 1, False, True),
                                                     (24, 24, 24, 1, False, True), (24, 0, 4, 0, False, False),
                                                     (8, 8, 8, 1, False, True), (8, 0, 3, 0, False, False),
                                                     (8, 0, 8, 1, True, False), (24, 0, 2, 0, False, False)])
@pytest.mark.parametrize('S,P,HW', [(1000, 1000, 1000), (3 * 64, 64, 58), (777, 777, 777),
                                    (5 * 1024, 1024, 1000), (4100, 4100, 4100)])  # the last two: tensor-core paths
def test_pwconv(cuda, ci1, ci2, co, act, res, bias, S, P, HW):
    from multimodal_3d_image_segmentation_b200 import ops
    g = torch.Generator().manual_seed(7)
    B = 2
    in1 = torch.randn(B, ci1, S, generator=g)
    in2 = torch.randn(B, ci2, S, generator=g) if ci2 else None
    if act:  # in1 plays the role of a SELU output in one of the backward variants
        in1 = F.selu(in1)
    w = torch.randn(co, ci1 + ci2, generator=g) / np.sqrt(ci1 + ci2)
    b = torch.randn(co, generator=g) * 0.1 if bias else None
    dy = torch.randn(B, co, S, generator=g)
    live = ((torch.arange(S) % P) < HW).double()
    r1 = in1.double().requires_grad_(True)
    r2 = in2.double().requires_grad_(True) if ci2 else None
    rw = w.double().requires_grad_(True)
    rb = b.double().requires_grad_(True) if bias else None
    yr = _pw_ref(r1, r2, rw, rb, act, res)
    (yr * dy.double() * live).sum().backward()
    dev = lambda t: None if t is None else t.to(cuda)  # noqa: E731
    y = ops.pwconv_forward(dev(in1), dev(in2), dev(w), dev(b), act, res)
    assert rel(y, yr) < 2e-6
    din1, din2, dw, db = ops.pwconv_backward(dev(dy), y, dev(in1), dev(in2), dev(w), act, res, hw=(P, HW),
                                             has_bias=bias)
    assert rel(din1, r1.grad) < 3e-6
    if ci2:
        assert rel(din2, r2.grad) < 3e-6
    assert rel(dw, rw.grad) < 3e-6
    if bias:
        assert rel(db, rb.grad) < 3e-6
    else:
        assert db is None
    # accumulate flags + "in1 is a SELU output" flag
    if act:
        base1 = torch.randn(B, ci1, S, generator=g).to(cuda)
        wacc = torch.ones(co, ci1 + ci2, device=cuda)
        d1, _, dw2, _ = ops.pwconv_backward(dev(dy), y, dev(in1), dev(in2), dev(w), act, res, hw=(P, HW),
                                            in1_is_selu=True, din1=base1.clone(), dweight=wacc, has_bias=bias,
                                            accumulate_w=True, need_in2=False)
        sg = torch.where(in1 > 0, torch.full_like(in1, orc.SELU_SCALE), in1 + orc.SELU_SCALE * orc.SELU_ALPHA).double()
        assert rel(d1, base1.cpu().double() + r1.grad * sg) < 3e-6
        assert rel(dw2, rw.grad + 1.0) < 3e-6


def test_hartley_conv_golden(cuda, golden_dir):
    from multimodal_3d_image_segmentation_b200 import ops
    g = dict(np.load(os.path.join(golden_dir, 'operator.npz')))
    z = torch.from_numpy(g['z']).to(cuda).requires_grad_(True)
    w = torch.from_numpy(g['w_individual']).to(cuda).requires_grad_(True)
    y = ops.HartleyConv.apply(z, w)
    assert maxrel(y, g['y_individual']) < 1e-5
    dz, dw = torch.autograd.grad(y, [z, w], torch.from_numpy(g['g_individual']).to(cuda))
    assert maxrel(dz, g['dz_individual']) < 1e-5 and maxrel(dw, g['dw_individual']) < 1e-5
    # fused residual + SELU variant against the oracle
    zr = torch.from_numpy(g['z']).double().requires_grad_(True)
    wr = torch.from_numpy(g['w_individual']).double().requires_grad_(True)
    yr = F.selu(orc.hartley_mix(zr, wr) + zr)
    gg = torch.from_numpy(g['g_individual'])
    dzr, dwr = torch.autograd.grad(yr, [zr, wr], gg.double())
    y2 = ops.HartleyConv.apply(z, w, True)
    dz2, dw2 = torch.autograd.grad(y2, [z, w], gg.to(cuda))
    assert rel(y2, yr) < 2e-6 and rel(dz2, dzr) < 3e-6 and rel(dw2, dwr) < 3e-6


# ---------------------------------------------------------------------------------------------- stem
@pytest.mark.parametrize('cin,f,shape', [(4, 24, (10, 9, 7)), (2, 8, (18, 16, 13)), (1, 8, (6, 6, 6)), (4, 24, (12, 12, 11))])
def test_stem(cuda, cin, f, shape):
    from multimodal_3d_image_segmentation_b200 import ops
    from multimodal_3d_image_segmentation_b200.plan import plane_pitch
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, cin, *shape, generator=g)
    w = torch.randn(f, cin, 2, 2, 2, generator=g) / np.sqrt(8 * cin)
    b = torch.randn(f, generator=g) * 0.1
    rw = w.double().requires_grad_(True)
    rb = b.double().requires_grad_(True)
    yr = F.selu(F.conv3d(x.double(), rw, rb, stride=2, padding=1))
    D, H, W = yr.shape[2:]
    assert (D, H, W) == ops.stem_out_shape(shape)
    dy = torch.randn(yr.shape, generator=g)
    (yr * dy.double()).sum().backward()
    for pitch in (None, plane_pitch(H, W)):
        y = ops.stem_forward(x.to(cuda), w.to(cuda), b.to(cuda), pitch=pitch)
        yd = from_planar(y, H, W) if pitch else y
        assert rel(yd, yr) < 2e-6
        if pitch and pitch > H * W:
            assert float(y[..., H * W:].abs().max()) == 0.0
        dpre = ops.selu_backward(dy.to(cuda), yd)
        dpre = to_planar(dpre, pitch) if pitch else dpre
        dw, db = ops.stem_backward(dpre, x.to(cuda), f, pitch=pitch)
        assert rel(dw, rw.grad) < 3e-6 and rel(db, rb.grad) < 3e-6


# ---------------------------------------------------------------------------------------------- head + losses
@pytest.mark.parametrize('lo,hi,C', [((10, 9, 7), (18, 16, 13), 3), ((6, 6, 6), (10, 10, 10), 4), ((7, 6, 5), (12, 11, 9), 2)])
def test_head_forward_backward(cuda, lo, hi, C):
    from multimodal_3d_image_segmentation_b200 import ops
    from multimodal_3d_image_segmentation_b200.plan import get_interp_tables, plane_pitch
    g = torch.Generator().manual_seed(9)
    ll = torch.randn(2, C, *lo, generator=g) * 3
    tables = get_interp_tables(lo, hi, cuda)
    lr = ll.double().requires_grad_(True)
    pr = torch.softmax(F.interpolate(lr, size=hi, mode='trilinear'), dim=1)
    dp = torch.randn(pr.shape, generator=g)
    (pr * dp.double()).sum().backward()
    H, W = lo[1:]
    for pitch in (H * W, plane_pitch(H, W)):
        lld = to_planar(ll.to(cuda), pitch) if pitch != H * W else ll.to(cuda)
        probs = ops.head_forward(lld, tables, pitch, 1)
        assert rel(probs, pr) < 3e-6
        dll = ops.head_backward(dp.to(cuda), probs, tables, pitch, 1)
        if pitch != H * W:
            assert float(dll[..., H * W:].abs().max()) == 0.0
            dll = from_planar(dll, H, W)
        assert rel(dll, lr.grad) < 5e-6
    # interpolation only (output_activation disabled)
    up = ops.head_forward(ll.to(cuda), tables, H * W, 0)
    assert rel(up, F.interpolate(ll.double(), size=hi, mode='trilinear')) < 2e-6


@pytest.mark.parametrize('name', ['DiceLoss', 'PCCLoss', 'ExpDiceLoss', 'CrossEntropyLoss'])
def test_losses_golden_and_oracle(cuda, golden_dir, name):
    from multimodal_3d_image_segmentation_b200 import nets
    g = dict(np.load(os.path.join(golden_dir, 'losses.npz')))
    p = torch.from_numpy(g['p']).to(cuda).requires_grad_(True)
    t = torch.from_numpy(g['t']).to(cuda)
    loss = getattr(nets.custom_losses, name)()(p, t)
    assert abs(float(loss) - float(g[f'{name}/loss'])) < 2e-6
    (grad,) = torch.autograd.grad(loss, p)
    assert maxrel(grad, g[f'{name}/grad']) < 2e-5
    # larger, ragged case vs the fp64 oracle
    gen = torch.Generator().manual_seed(10)
    pp = torch.softmax(torch.randn(2, 4, 33, 31, 29, generator=gen), 1)
    tt = orc.to_categorical(torch.randint(0, 4, (2, 1, 33, 31, 29), generator=gen), 4)
    pr = pp.double().requires_grad_(True)
    lr = orc.LOSSES[name](pr, tt.double())
    (gr,) = torch.autograd.grad(lr, pr)
    pc = pp.to(cuda).requires_grad_(True)
    lc = getattr(nets.custom_losses, name)()(pc, tt.to(cuda))
    (gc,) = torch.autograd.grad(lc * 3.0, pc)
    assert abs(float(lc) - float(lr)) < 1e-6 and rel(gc, 3.0 * gr) < 1e-5


def test_exp_dice_exponent_and_ce_arguments(cuda):
    from multimodal_3d_image_segmentation_b200 import nets
    gen = torch.Generator().manual_seed(13)
    pp = torch.softmax(torch.randn(1, 3, 9, 8, 7, generator=gen), 1)
    tt = orc.to_categorical(torch.randint(0, 3, (1, 1, 9, 8, 7), generator=gen), 3)
    for e in (0.3, 1.0, 2.5):
        pr = pp.double().requires_grad_(True)
        lr = orc.exp_dice_loss(pr, tt.double(), e)
        (gr,) = torch.autograd.grad(lr, pr)
        pc = pp.to(cuda).requires_grad_(True)
        lc = nets.custom_losses.ExpDiceLoss(exp=e)(pc, tt.to(cuda))
        (gc,) = torch.autograd.grad(lc, pc)
        assert abs(float(lc) - float(lr)) < 2e-6 * max(1.0, float(lr)) and rel(gc, gr) < 1e-5, e
    with pytest.raises(ValueError):
        nets.custom_losses.ExpDiceLoss(exp=0.0)
    with pytest.raises(NotImplementedError):
        nets.custom_losses.CrossEntropyLoss(label_smoothing=0.1)
    with pytest.raises(ValueError):  # class-index targets are not what the reference passes (train_test.py:152)
        nets.custom_losses.CrossEntropyLoss()(pp.to(cuda), torch.zeros(1, 9, 8, 7, device=cuda))


@pytest.mark.parametrize('shape', [(2, 4, 33, 31, 29), (1, 3, 5, 4, 3), (3, 1, 17, 16, 15), (1, 8, 40, 9, 11)])
def test_cross_entropy_labels_soft_targets_and_torch(cuda, shape):
    """The CE kernels on uint8 labels equal the one-hot route bit for bit; soft (non one-hot) class-probability targets
    and unnormalised inputs follow torch.nn.CrossEntropyLoss (the class the reference instantiates, run.py:110)."""
    from multimodal_3d_image_segmentation_b200 import nets, ops
    B, C = shape[:2]
    gen = torch.Generator().manual_seed(14)
    pp = torch.softmax(torch.randn(*shape, generator=gen), 1)
    labels = torch.randint(0, C, (B, 1) + shape[2:], generator=gen)
    onehot = orc.to_categorical(labels, C).contiguous()  # (the raw entry points take dense NCDHW tensors)
    pc = pp.to(cuda)
    lab = labels[:, 0].to(torch.uint8).to(cuda).contiguous()
    if C > 1:  # to_categorical returns a channels-last view, as the reference's moveaxis does (utils.py:96)
        with pytest.raises(ValueError):
            ops.ce_loss_forward(pc, y_true=orc.to_categorical(labels, C).to(cuda))
    l_lab = ops.ce_loss_forward(pc, labels=lab)
    l_hot = ops.ce_loss_forward(pc, y_true=onehot.to(cuda))
    assert float(l_lab) == float(l_hot)
    gscale = torch.tensor([0.37], device=cuda)
    g_lab = ops.ce_loss_backward(pc, labels=lab, grad_loss=gscale)
    g_hot = ops.ce_loss_backward(pc, y_true=onehot.to(cuda), grad_loss=gscale)
    assert torch.equal(g_lab, g_hot)
    pr = pp.double().requires_grad_(True)
    lr = orc.cross_entropy_loss(pr, onehot.double())
    (gr,) = torch.autograd.grad(lr, pr)
    assert abs(float(l_lab) - float(lr)) < 1e-6 and rel(g_lab, 0.37 * gr) < 1e-5
    # soft targets, raw scores instead of probabilities
    xs = torch.randn(*shape, generator=gen) * 3
    ts = torch.softmax(torch.randn(*shape, generator=gen), 1)
    xr = xs.double().requires_grad_(True)
    lt = torch.nn.CrossEntropyLoss()(xr, ts.double())
    (gt,) = torch.autograd.grad(lt, xr)
    xc = xs.to(cuda).requires_grad_(True)
    lc = nets.custom_losses.CrossEntropyLoss()(xc, ts.to(cuda))
    (gc,) = torch.autograd.grad(lc, xc)
    assert abs(float(lc) - float(lt)) < 2e-6 * max(1.0, float(lt)) and rel(gc, gt) < 1e-5


def test_losses_full_size_known_answers(cuda):
    """BASELINE volume size (2 x 4 x 240 x 240 x 155): closed-form values that do not depend on the size.
    Perfect predictions p = one-hot(t): Dice = 1 (loss 0), Pearson r = 1 (PCC loss 0), ExpDice clamps at 1 - 1e-7,
    cross entropy = log(e + C - 1) - 1; uniform predictions p = 1/C: cross entropy = log C, and its gradient sums to 0
    over the classes of every voxel."""
    import math
    from multimodal_3d_image_segmentation_b200 import nets, ops
    from multimodal_3d_image_segmentation_b200.experiments import to_categorical
    B, C, shape = 2, 4, (240, 240, 155)
    lab = torch.randint(0, C, (B, 1) + shape, device=cuda, dtype=torch.uint8,
                        generator=torch.Generator(device=cuda).manual_seed(20))
    t = to_categorical(lab, C)
    assert float(t.sum(dtype=torch.float64)) == B * shape[0] * shape[1] * shape[2]
    assert abs(float(nets.custom_losses.DiceLoss()(t, t))) < 1e-6
    assert abs(float(nets.custom_losses.PCCLoss()(t, t))) < 1e-6
    assert abs(float(nets.custom_losses.ExpDiceLoss()(t, t)) - (-math.log(1 - 1e-7)) ** 0.3) < 1e-3
    ce = nets.custom_losses.CrossEntropyLoss()
    assert abs(float(ce(t, t)) - (math.log(math.e + C - 1) - 1.0)) < 1e-6
    u = torch.full_like(t, 1.0 / C).requires_grad_(True)
    loss = ce(u, t)
    assert abs(float(loss) - math.log(C)) < 1e-6
    (g,) = torch.autograd.grad(loss, u)
    assert float(g.sum(1).abs().max()) < 1e-12 and abs(float(g.abs().sum()) - 2.0 * (C - 1) / C) < 1e-4
    assert float(ops.ce_loss_forward(u.detach(), labels=lab[:, 0].contiguous())) == float(loss)


@pytest.mark.parametrize('kind', ['DiceLoss', 'PCCLoss', 'ExpDiceLoss'])
def test_fused_head_loss(cuda, kind):
    from multimodal_3d_image_segmentation_b200 import ops
    from multimodal_3d_image_segmentation_b200.plan import get_interp_tables, plane_pitch
    lo, hi, C = (10, 9, 7), (18, 16, 13), 3
    g = torch.Generator().manual_seed(11)
    ll = torch.randn(2, C, *lo, generator=g) * 2
    labels = torch.randint(0, C, (2, 1, *hi), generator=g)
    lr = ll.double().requires_grad_(True)
    pr = torch.softmax(F.interpolate(lr, size=hi, mode='trilinear'), dim=1)
    loss_r = orc.LOSSES[kind](pr, orc.to_categorical(labels, C).double())
    loss_r.backward()
    tables = get_interp_tables(lo, hi, cuda)
    pitch = plane_pitch(lo[1], lo[2])
    lld = to_planar(ll.to(cuda), pitch)
    lab = labels[:, 0].to(torch.uint8).to(cuda).contiguous()
    loss, coef = ops.head_loss_forward(lld, lab, tables, pitch, ops.LOSS_KINDS[kind], ops.LOSS_DEFAULT_PARAM[kind])
    assert abs(float(loss) - float(loss_r)) < 1e-6
    dll = ops.head_loss_backward(lld, lab, coef, None, tables, pitch)
    assert rel(from_planar(dll, lo[1], lo[2]), lr.grad) < 1e-5


def test_adamax(cuda):
    from multimodal_3d_image_segmentation_b200 import _lib
    g = torch.Generator().manual_seed(12)
    p0 = torch.randn(1000, generator=g)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adamax([pr], lr=5e-3)
    pc = p0.to(cuda)
    m = torch.zeros_like(pc)
    u = torch.zeros_like(pc)
    for step in range(1, 4):
        grad = torch.randn(1000, generator=g)
        pr.grad = grad.clone()
        opt.step()
        _lib.call('hno_adamax_step', pc.data_ptr(), grad.to(cuda).data_ptr(), m.data_ptr(), u.data_ptr(), 1000, 5e-3,
                  0.9, 0.999, 1e-8, 0.0, step, 1.0, torch.cuda.current_stream().cuda_stream)
    assert maxrel(pc, pr) < 1e-6


@pytest.mark.parametrize('C,L,M,B', [(24, 3, 20 * 28 * 28, 2), (24, 1, 777, 1), (8, 4, 1000, 3)])
def test_modechain_matches_layerwise(cuda, C, L, M, B):
    """The fused n_XS-mix kernel against the oracle's per-layer selu(W z + z) chain and its autograd (fp64)."""
    import torch
    from multimodal_3d_image_segmentation_b200 import ops
    g = torch.Generator().manual_seed(5)
    z0 = torch.randn(B, C, M, generator=g)
    ws = [torch.randn(C, C, generator=g) / C ** 0.5 for _ in range(L)]
    dz = torch.randn(B, C, M, generator=g)
    zr = z0.double().requires_grad_(True)
    wr = [w.double().requires_grad_(True) for w in ws]
    cur = zr
    outs = []
    for w in wr:
        cur = torch.nn.functional.selu(torch.einsum('oi,bim->bom', w, cur) + cur)
        outs.append(cur)
    grads = torch.autograd.grad(cur, [zr] + wr, dz.double())
    zs = ops.modechain_forward(z0.to(cuda), [w.to(cuda) for w in ws])
    for l in range(L):
        assert torch.allclose(zs[l].cpu().double(), outs[l].detach(), rtol=2e-5, atol=2e-5)
    dz0, dws = ops.modechain_backward(dz.to(cuda), z0.to(cuda), zs, [w.to(cuda) for w in ws])
    assert ((dz0.cpu().double() - grads[0]).norm() / grads[0].norm()).item() < 1e-5
    for l in range(L):
        assert ((dws[l].cpu().double() - grads[1 + l]).norm() / grads[1 + l].norm()).item() < 1e-5


# ---------------------------------------------------------------------------------------------- input side (8f-4)
def test_to_categorical(cuda, golden_dir):
    from multimodal_3d_image_segmentation_b200.experiments import to_categorical
    g = dict(np.load(os.path.join(golden_dir, 'input_side.npz')))
    for dt in (torch.uint8, torch.int64, torch.int32, torch.float32):  # train_test.py feeds whatever the loader yields
        y = to_categorical(torch.from_numpy(g['labels']).to(cuda).to(dt), 4)
        assert y.dtype == torch.float32 and y.is_contiguous() and np.array_equal(y.cpu().numpy(), g['onehot'])
    gen = torch.Generator().manual_seed(18)
    for shape, C in (((2, 1, 24, 20, 31), 4), ((1, 1, 5, 3, 7), 2), ((3, 1, 16, 16, 16), 7)):  # N % 4 != 0 and == 0
        lab = torch.randint(0, C, shape, generator=gen)
        ref = orc.to_categorical(lab, C)
        for dt in (torch.uint8, torch.int64):
            assert torch.equal(to_categorical(lab.to(dt).to(cuda), C).cpu(), ref)
        assert torch.equal(to_categorical(lab.to(torch.uint8).to(cuda)).cpu(), orc.to_categorical(lab, int(lab.max()) + 1))
    with pytest.raises(IndexError):  # the reference's scatter raises on labels >= num_classes
        to_categorical(torch.full((1, 1, 4, 4, 4), 5, dtype=torch.uint8, device=cuda), 4)
    with pytest.raises(IndexError):
        to_categorical(torch.full((1, 1, 4, 4, 3), -1, dtype=torch.int64, device=cuda), 4)
    with pytest.raises(RuntimeError):
        to_categorical(torch.zeros(1, 1, 4, 4, 4, dtype=torch.uint8), 4)
    # feeds the drop-in losses exactly like train_test.py:152-160
    from multimodal_3d_image_segmentation_b200 import nets
    p = torch.softmax(torch.randn(2, 4, 24, 20, 31, generator=gen), 1)
    lab = torch.randint(0, 4, (2, 1, 24, 20, 31), generator=gen)
    loss = nets.custom_losses.PCCLoss()(p.to(cuda), to_categorical(lab.to(cuda), 4))
    assert abs(float(loss) - float(orc.pcc_loss(p.double(), orc.to_categorical(lab, 4).double()))) < 1e-6


def test_normalize_modalities(cuda, golden_dir):
    from multimodal_3d_image_segmentation_b200.experiments import normalize_modalities
    g = dict(np.load(os.path.join(golden_dir, 'input_side.npz')))
    cases = {'plain': {}, 'mask': dict(mask_val=0), 'clip': dict(clip_val=(50.0, 900.0)),
             'maskclip': dict(mask_val=0, clip_val=(0.0, 700.0)), 'maskhit': dict(mask_val=700, clip_val=(0.0, 700.0))}
    vol = torch.from_numpy(g['vol']).to(cuda)
    for tag, kw in cases.items():
        y = normalize_modalities(vol, **kw)
        assert y.dtype == torch.float32 and float((y.cpu() - torch.from_numpy(g[f'norm/{tag}'])).abs().max()) < 1e-5, tag
    # BraTS-shaped volume (4 x 155 x 240 x 240, background 0) against the numpy oracle; int16 input as SimpleITK yields
    rng = np.random.default_rng(19)
    big = rng.gamma(4.0, 150.0, (4, 155, 240, 240)).astype(np.float32)
    big *= (rng.random((1, 155, 240, 240)) < 0.4)
    big = np.round(big).astype(np.int16)
    ref = orc.normalize_modalities(big, mask_val=0)
    y = normalize_modalities(torch.from_numpy(big).to(cuda), mask_val=0).cpu().numpy()
    assert np.abs(y - ref).max() < 2e-5
    assert np.all(y[big == 0] == 0) and abs(float(y[0][big[0] != 0].mean())) < 1e-4
    # ragged length (n % 4 != 0), an all-background modality
    odd = rng.normal(100.0, 20.0, (3, 5, 7, 9)).astype(np.float32)
    odd[1] = 0
    y = normalize_modalities(torch.from_numpy(odd).to(cuda), mask_val=0).cpu().numpy()
    assert np.abs(y - orc.normalize_modalities(odd, mask_val=0)).max() < 1e-5 and np.all(y[1] == 0)
    # the int16 kernels (raw NIfTI storage type on the wire) are bit-identical to the fp32 ones, also at ragged lengths and
    # for a batch normalised row by row into a caller-provided buffer
    from multimodal_3d_image_segmentation_b200.experiments.utils import normalize_rows
    yf = normalize_modalities(torch.from_numpy(big.astype(np.float32)).to(cuda), mask_val=0).cpu().numpy()
    assert np.array_equal(y16 := normalize_modalities(torch.from_numpy(big).to(cuda), mask_val=0).cpu().numpy(), yf)
    odd16 = np.round(odd * 3).astype(np.int16)
    a = normalize_modalities(torch.from_numpy(odd16).to(cuda), mask_val=0).cpu().numpy()
    b = normalize_modalities(torch.from_numpy(odd16.astype(np.float32)).to(cuda), mask_val=0).cpu().numpy()
    assert np.array_equal(a, b)
    batch = torch.from_numpy(np.stack([big[:, :20], big[:, 20:40]])).to(cuda)  # (2, 4, 20, 240, 240) int16
    out = torch.empty(batch.shape, dtype=torch.float32, device=cuda)
    normalize_rows(batch, 8, mask_val=0, out=out)
    for i in range(2):
        assert np.abs(out[i].cpu().numpy() - orc.normalize_modalities(batch[i].cpu().numpy(), mask_val=0)).max() < 2e-5
    del y16


@pytest.mark.parametrize('grid,modes', [((25, 33, 32), (10, 14, 14)), ((24, 40, 28), (12, 14, 14)), ((121, 121, 78), (10, 14, 14))])
def test_fused_spectral_chain_matches_separate_kernels(cuda, grid, modes):
    """hno_dht3_chain_forward / _backward (transform -> n_XS shared-weight mixes -> inverse transform with the W stages, the cas
    recombination and the mixes in ONE kernel, csrc/spectral_core.cu) against the separate launches it replaces
    (hno_dht3_forward -> hno_modechain_* -> hno_dht3_adjoint), incl. a grid whose retained sets contain the Nyquist frequency
    (n == 2m: sine rows vanish) and the BASELINE grid."""
    from multimodal_3d_image_segmentation_b200 import ops
    from multimodal_3d_image_segmentation_b200.plan import get_crop_plan, plane_pitch
    D, H, W = grid
    P = plane_pitch(H, W)
    B, C, L = 2, 24, 3
    plan = get_crop_plan(grid, modes, cuda)
    g = torch.Generator(device=cuda).manual_seed(12)
    x = torch.randn(B, C, D, P, device=cuda, generator=g)
    x.view(B, C, D, P)[..., H * W:] = 0
    ws = [torch.randn(C, C, device=cuda, generator=g) * 0.2 for _ in range(L)]
    inv_n = 1.0 / plan.n_voxels
    assert ops.dht3_chain_eligible(x, plan, C, L)
    # forward
    z0 = ops.dht3_forward(x, plan, inv_n)
    zs = ops.modechain_forward(z0, ws)
    u_ref = ops.dht3_adjoint(zs[-1], plan, 1.0, epilogue=2, pitch=P)
    u, zall = ops.dht3_chain_forward(x, plan, ws, inv_n, epilogue=2, save=True)
    valid = torch.zeros(P, dtype=torch.bool, device=cuda)
    valid[:H * W] = True
    assert rel(zall[0], z0) < 2e-6, rel(zall[0], z0)
    for l in range(L):
        assert rel(zall[l + 1], zs[l]) < 5e-6, (l, rel(zall[l + 1], zs[l]))
    assert rel(u[..., valid], u_ref[..., valid]) < 5e-6, rel(u[..., valid], u_ref[..., valid])
    # backward: dxin += (1/N) C^T chain_bwd(C dt)
    dt = torch.randn(B, C, D, P, device=cuda, generator=g)
    dt[..., H * W:] = 0
    base = torch.randn(B, C, D, P, device=cuda, generator=g)
    dz = ops.dht3_forward(dt, plan, 1.0)
    dz0, dws_ref = ops.modechain_backward(dz, z0, zs, ws)
    ref = base.clone()
    ops.dht3_adjoint(dz0, plan, inv_n, epilogue=1, out=ref)
    out = base.clone()
    dws = ops.dht3_chain_backward(dt, plan, zall, ws, inv_n, out, epilogue=1)
    assert rel(out[..., valid], ref[..., valid]) < 5e-6, rel(out[..., valid], ref[..., valid])
    for l in range(L):
        assert rel(dws[l], dws_ref[l]) < 2e-5, (l, rel(dws[l], dws_ref[l]))


@pytest.mark.parametrize('wt', ['shared', 'individual'])
@pytest.mark.parametrize('ci,co,shape,modes', [(4, 8, (12, 10, 9), (3, 2, 4)), (8, 4, (8, 8, 8), (4, 4, 4)),
                                               (24, 24, (20, 18, 14), (6, 5, 7))])
def test_fourier_mix_matches_eager_formula(cuda, ci, co, shape, modes, wt):
    """hno_fourier_mix_* (Re / Im split over the (k, N - k) pairs, complex channel mix with shared or per-mode weights, c_k
    re-assembly; the mode-domain step of nets/fourier_operator.py:155, 165-209) against the same formula in fp64 torch,
    forward and all gradients, incl. grids whose retained sets contain self-conjugate modes (n == 2 m)."""
    from multimodal_3d_image_segmentation_b200 import nets, ops
    op = nets.FourierOperator(ci, co, modes, weights_type=wt, device=cuda)
    plan, lin_k, lin_n, kshape, ls, ck, fused = op._geometry(shape, cuda)
    MS = ls[0] * ls[1] * ls[2]
    MK = lin_k.numel()
    g = torch.Generator().manual_seed(11)
    z = torch.randn(2, ci, MS, generator=g)
    wshape = (co, ci) if wt == 'shared' else (co, ci) + tuple(kshape)
    wr = torch.randn(wshape, generator=g) * 0.3
    wi = torch.randn(wshape, generator=g) * 0.3
    gy = torch.randn(2, co, MS, generator=g)
    # fp64 reference
    zr, wrr, wir = (t.double().requires_grad_(True) for t in (z, wr, wi))
    lk, ln = lin_k.cpu(), lin_n.cpu()
    hk, hn = zr.index_select(2, lk), zr.index_select(2, ln)
    re, im = (hk + hn) * 0.5, (hn - hk) * 0.5
    eq = 'oi,bim->bom' if wt == 'shared' else 'oim,bim->bom'
    w2r, w2i = (wrr, wir) if wt == 'shared' else (wrr.reshape(co, ci, MK), wir.reshape(co, ci, MK))
    a = torch.einsum(eq, w2r, re) - torch.einsum(eq, w2i, im)
    b = torch.einsum(eq, w2i, re) + torch.einsum(eq, w2r, im)
    ckf = fused[2].cpu().double().view(1, 1, -1)
    hp = torch.zeros(2, co, MS, dtype=torch.float64)
    hp = hp.index_add(2, lk, ckf * (a - b) * 0.5).index_add(2, ln, ckf * (a + b) * 0.5)
    (hp * gy.double()).sum().backward()
    # CUDA
    zc, wrc, wic = (t.to(cuda).requires_grad_(True) for t in (z, wr, wi))
    out = ops.FourierMixShared.apply(zc, wrc, wic, *fused)
    assert rel(out, hp) < 2e-6
    (out * gy.to(cuda)).sum().backward()
    assert rel(zc.grad, zr.grad) < 2e-6
    assert rel(wrc.grad, wrr.grad) < 2e-6 and rel(wic.grad, wir.grad) < 2e-6
    out2 = ops.FourierMixShared.apply(zc, wrc, wic, *fused)  # two-term atomic sums: bit-identical from run to run
    assert torch.equal(out, out2)


@pytest.mark.parametrize('dtype', [torch.float32, torch.int16, torch.uint8])
def test_permute_spatial_is_a_pure_layout_change(cuda, dtype):
    """hno_transpose2d behind ops.permute_spatial: both 'one axis to the end' permutations, ragged sizes, every element size."""
    from multimodal_3d_image_segmentation_b200 import ops as hops
    gen = torch.Generator().manual_seed(8)
    for shape in ((2, 3, 5, 33, 70), (1, 1, 37, 31, 1), (3, 40, 64, 9)):  # (..., D, H, W)
        t = torch.randint(0, 200, shape, generator=gen).to(dtype).to(cuda)
        lead = len(shape) - 3
        for perm in ((1, 2, 0), (0, 2, 1), (2, 0, 1)):
            want = t.permute(*range(lead), *[lead + p for p in perm]).contiguous()
            got = hops.permute_spatial(t, perm)
            assert got.shape == want.shape and torch.equal(got, want), (shape, perm)
