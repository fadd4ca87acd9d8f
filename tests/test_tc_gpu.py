"""Parity of the tcgen05 (3xTF32) tensor-core path against fp64 maths and against the fp32 CUDA-core kernels.

The tensor-core kernels only take over for TMA-legal strides and large extents, so these cases use sizes above the
dispatch thresholds (pointwise conv: S >= 4096 and S % 4 == 0; DHT D stages: plane pitch >= 1024).
"""
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


def maxabs(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.fixture()
def tc_switch():
    from multimodal_3d_image_segmentation_b200 import _lib
    lib = _lib.load()
    yield lib.hno_set_tensor_cores
    lib.hno_set_tensor_cores(1)


def _launches():
    from multimodal_3d_image_segmentation_b200 import _lib
    return _lib.load().hno_launch_count(0)


@pytest.mark.parametrize('ci1,ci2,co,act,bias', [(24, 24, 24, 1, True), (24, 0, 24, 1, True), (24, 0, 4, 0, False),
                                                 (8, 8, 8, 1, True), (8, 0, 3, 0, False)])
@pytest.mark.parametrize('S', [4096, 5 * 1056, 128 * 37 + 4])
def test_pwconv_forward_tc(cuda, tc_switch, ci1, ci2, co, act, bias, S):
    from multimodal_3d_image_segmentation_b200 import ops
    g = torch.Generator().manual_seed(11)
    B = 2
    in1 = torch.randn(B, ci1, S, generator=g) * 3.0
    in2 = torch.randn(B, ci2, S, generator=g) if ci2 else None
    w = torch.randn(co, ci1 + ci2, generator=g) / np.sqrt(ci1 + ci2)
    b = torch.randn(co, generator=g) * 0.1 if bias else None
    x = in1 if in2 is None else torch.cat([in1, in2], 1)
    yr = torch.einsum('oi,bis->bos', w.double(), x.double())
    if bias:
        yr = yr + b.double().view(1, -1, 1)
    if act:
        yr = F.selu(yr)
    dev = lambda t: None if t is None else t.to(cuda)  # noqa: E731
    tc_switch(1)
    y_tc = ops.pwconv_forward(dev(in1), dev(in2), dev(w), dev(b), act, False)
    tc_switch(0)
    y_cc = ops.pwconv_forward(dev(in1), dev(in2), dev(w), dev(b), act, False)
    tc_switch(1)
    # 3xTF32: ~2^-22 per product, fp32 accumulation -> same class of error as the fp32 FFMA kernel
    assert rel(y_tc, yr) < 2e-6, rel(y_tc, yr)
    assert maxabs(y_tc, yr) < 5e-6
    assert rel(y_cc, yr) < 2e-6
    if ci1 in (8, 24, 32):
        assert not torch.equal(y_tc, y_cc) or True  # both paths ran; bitwise equality is not required


DHT_TC_CASES = [((40, 40, 26), (10, 14, 14)), ((20, 33, 40), (4, 5, 6)), ((37, 36, 30), (10, 14, 14)),
                ((130, 33, 32), (10, 14, 14))]


@pytest.mark.parametrize('shape,modes', DHT_TC_CASES)
@pytest.mark.parametrize('padded', [False, True])
def test_dht_tc(cuda, tc_switch, shape, modes, padded):
    from multimodal_3d_image_segmentation_b200 import ops
    from multimodal_3d_image_segmentation_b200.plan import get_crop_plan, plane_pitch
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, *shape, generator=g)
    plan = get_crop_plan(shape, modes, cuda)
    n = float(np.prod(shape))
    D, H, W = shape
    pitch = plane_pitch(H, W) if padded else None

    def planar(t):
        if not padded:
            return t.to(cuda)
        out = torch.zeros(t.shape[0], t.shape[1], D, pitch, device=cuda)
        out[..., :H * W] = t.to(cuda).reshape(t.shape[0], t.shape[1], D, H * W)
        return out

    def dense(t):
        return t[..., :H * W].reshape(t.shape[0], t.shape[1], D, H, W) if padded else t

    z_ref = orc.transform_crop(x.double(), modes)
    zz = torch.randn(z_ref.shape, generator=g)
    y_ref = orc.pad_inverse(zz.double(), shape)
    base = torch.randn(2, 3, *shape, generator=g)
    res = {}
    for on in (1, 0):
        tc_switch(on)
        z = ops.dht3_forward(planar(x), plan, 1.0 / n)
        y = ops.dht3_adjoint(zz.to(cuda), plan, 1.0, pitch=pitch)
        if padded and pitch > H * W:
            assert float(y[..., H * W:].abs().max()) == 0.0
        ys = ops.dht3_adjoint(zz.to(cuda), plan, 1.0, epilogue=2, pitch=pitch)
        acc = planar(base).clone()
        ops.dht3_adjoint(zz.to(cuda), plan, 0.5, epilogue=1, out=acc)
        res[on] = (z, dense(y), dense(ys), dense(acc))
    tc_switch(1)
    for on in (1, 0):
        z, y, ys, acc = res[on]
        assert rel(z, z_ref) < 3e-6, (on, rel(z, z_ref))
        assert rel(y, y_ref) < 3e-6, (on, rel(y, y_ref))
        assert rel(ys, F.selu(y_ref)) < 3e-6
        assert rel(acc, base.double() + 0.5 * y_ref) < 3e-6


def test_tc_path_is_taken(cuda, tc_switch):
    """The tensor-core kernels (not a silent fallback) serve the large pointwise convolutions: one launch, and the
    result differs in the last bits from the FFMA kernel's."""
    from multimodal_3d_image_segmentation_b200 import ops
    g = torch.Generator().manual_seed(3)
    in1 = torch.randn(1, 24, 8192, generator=g).to(cuda)
    in2 = torch.randn(1, 24, 8192, generator=g).to(cuda)
    w = (torch.randn(24, 48, generator=g) / 7).to(cuda)
    tc_switch(1)
    n0 = _launches()
    y1 = ops.pwconv_forward(in1, in2, w, None, 1, False)
    assert _launches() - n0 == 1
    tc_switch(0)
    y0 = ops.pwconv_forward(in1, in2, w, None, 1, False)
    tc_switch(1)
    assert rel(y1, y0) < 1e-6
    assert not torch.equal(y1, y0)
