"""CPU-only checks of the host logic: DHT plan tables + decomposition, interpolation tables, C ABI surface."""
import os
import re

import numpy as np
import pytest
import torch

from multimodal_3d_image_segmentation_b200 import _lib
from multimodal_3d_image_segmentation_b200.plan import (corner_frequencies, get_crop_plan, get_dht_plan,
                                                       get_interp_tables, plane_pitch)
from oracle import hno_oracle as orc
from tests import emulate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'hno_b200.h')).read()
    declared = set(re.findall(r'\b(hno_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} is declared in include/hno_b200.h but not exported'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.hno_version() == 100


def test_errors_are_reported_not_thrown():
    lib = _lib.load()
    rc = lib.hno_dht3_plan_fill(None, 0, 4, 4, 4, None, 0, None, 0, None, 0)
    assert rc != 0 and 'null' in _lib.last_error()
    with pytest.raises(_lib.HnoError):
        _lib.call('hno_pwconv_forward', None, None, None, None, None, 1, 24, 0, 24, 16, 1, 0, None)


def test_corner_frequencies_clamp():
    assert corner_frequencies(9, 2) == [0, 1, 7, 8]
    assert corner_frequencies(12, 10) == [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]  # 2m > n -> m = n // 2
    assert corner_frequencies(9, 14) == [0, 1, 2, 3, 5, 6, 7, 8]
    assert plane_pitch(121, 78) == 9440


@pytest.mark.parametrize('shape,modes', [((9, 8, 7), (2, 3, 3)), ((12, 10, 9), (10, 14, 14)), ((16, 11, 10), (4, 3, 5)),
                                         ((6, 6, 6), (3, 3, 3)), ((5, 4, 3), (1, 1, 1))])
def test_plan_decomposition_matches_oracle(shape, modes):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2,) + shape)
    plan = get_crop_plan(shape, modes, None)
    n = float(np.prod(shape))
    z = emulate.forward(plan, x, 1.0 / n)
    z_ref = orc.transform_crop(torch.from_numpy(x)[None], modes)[0].numpy()  # fp64 FFT route
    assert z.shape == z_ref.shape
    np.testing.assert_allclose(z, z_ref, atol=2e-7 * np.abs(z_ref).max(), rtol=0)
    z_dense = orc.dht3_dense(x, plan.klists, 1.0 / n)
    np.testing.assert_allclose(z, z_dense, atol=2e-7 * np.abs(z_dense).max(), rtol=0)
    zz = rng.standard_normal(z.shape)
    y = emulate.adjoint(plan, zz, 1.0)
    y_ref = orc.pad_inverse(torch.from_numpy(zz)[None], shape)[0].numpy()
    np.testing.assert_allclose(y, y_ref, atol=2e-7 * np.abs(y_ref).max(), rtol=0)
    # adjointness <C x, z> == <x, C^T z>
    lhs = float((emulate.forward(plan, x, 1.0) * zz).sum())
    rhs = float((x * y).sum())
    assert abs(lhs - rhs) <= 1e-9 * max(abs(lhs), 1.0)


def test_plan_full_spectrum_is_dhtn():
    shape = (5, 6, 4)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(shape)
    plan = get_dht_plan(shape, [list(range(n)) for n in shape], None)
    z = emulate.forward(plan, x, 1.0 / np.prod(shape))
    ref = orc.dhtn(torch.from_numpy(x)).numpy()
    np.testing.assert_allclose(z, ref, atol=1e-7)
    # involution: inverse(forward(x)) == x
    back = emulate.adjoint(plan, z, 1.0)
    np.testing.assert_allclose(back, x, atol=2e-6)


def test_baseline_plan_geometry():
    plan = get_crop_plan((121, 121, 78), (10, 14, 14), None)
    assert plan.modes_shape == (20, 28, 28)
    assert [(a['JC'], a['JS']) for a in plan.axes] == [(11, 10), (15, 14), (15, 14)]
    full = plan.table(0, 'full')
    # rows are exact cos / sin samples generated in fp64
    i = np.arange(121)
    np.testing.assert_allclose(full[3], np.cos(2 * np.pi * 3 * i / 121), atol=1e-7)
    np.testing.assert_allclose(full[11 + 2], np.sin(2 * np.pi * 3 * i / 121), atol=1e-7)


@pytest.mark.parametrize('lo,hi', [((10, 9, 7), (18, 16, 13)), ((121, 121, 78), (240, 240, 155)), ((5, 5, 5), (5, 5, 5)),
                                   ((7, 6, 5), (12, 11, 9))])
def test_interp_tables_match_torch_interpolate(lo, hi):
    t = get_interp_tables(lo, hi, None)
    if max(hi) > 100:  # big case: check the per-axis operators on 1-D ramps instead of a dense 3-D volume
        for a in range(3):
            M = emulate.interp_matrix(t, a)
            v = torch.randn(1, 1, lo[a], 1, 1, generator=torch.Generator().manual_seed(a))
            ref = torch.nn.functional.interpolate(v, size=(hi[a], 1, 1), mode='trilinear')[0, 0, :, 0, 0].numpy()
            np.testing.assert_allclose(M @ v[0, 0, :, 0, 0].numpy().astype(np.float64), ref, atol=2e-6)
        return
    x = torch.randn(1, 2, *lo, generator=torch.Generator().manual_seed(3))
    ref = torch.nn.functional.interpolate(x, size=hi, mode='trilinear').numpy()
    Md, Mh, Mw = (emulate.interp_matrix(t, a) for a in range(3))
    got = np.einsum('ad,bh,cw,ncdhw->ncabc'.replace('abc', 'xyz').replace('ad', 'xd').replace('bh', 'yh').replace('cw', 'zw'),
                    Md, Mh, Mw, x.numpy().astype(np.float64))
    np.testing.assert_allclose(got, ref, atol=2e-6)
    # inverse ranges cover exactly the touching indices
    for a in range(3):
        i0, i1, l1, s, e = t.axis(a)
        for lo_idx in range(lo[a]):
            touching = [o for o in range(hi[a]) if i0[o] == lo_idx or i1[o] == lo_idx]
            if touching:
                assert s[lo_idx] == touching[0] and e[lo_idx] == touching[-1] + 1


def test_stem_input_gradient_mapping_is_the_transposed_convolution():
    """The tap / voxel mapping k_stem_dx uses == autograd of Conv3d(kernel 2, stride 2, padding 1) w.r.t. its input."""
    import torch
    torch.manual_seed(0)
    for image in ((5, 6, 7), (4, 4, 4), (7, 3, 2)):
        x = torch.randn(2, 3, *image, dtype=torch.float64, requires_grad=True)
        wgt = torch.randn(5, 3, 2, 2, 2, dtype=torch.float64)
        y = torch.nn.functional.conv3d(x, wgt, stride=2, padding=1)
        g = torch.randn_like(y)
        (dx,) = torch.autograd.grad((y * g).sum(), x)
        got = emulate.stem_input_gradient(g.numpy(), wgt.numpy(), image)
        assert np.abs(got - dx.numpy()).max() < 1e-12, image


def test_resample_index_walk_matches_unravel():
    for (D, H, W) in ((3, 4, 8), (5, 3, 4), (2, 6, 2), (4, 1, 3)):
        N = D * H * W
        if N % 4:
            continue
        walk = emulate.resample_index_walk(N, W, H)
        want = np.stack(np.unravel_index(np.arange(N), (D, H, W)), 1)
        assert np.array_equal(walk, want), (D, H, W)
