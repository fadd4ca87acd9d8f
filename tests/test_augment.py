"""Affine augmentation (reference experiments/data_io/dataset.py:63-245).

CPU: the oracle's restatement and the package's host-side parameter drawing replayed against tests/golden/augment.npz,
which oracle/make_golden.py::case_augment recorded from the REFERENCE's ImageTransform (matrices handed to
sitk.AffineTransform, flips, augmented arrays).  GPU: `hno_affine_resample_nn` bit-exact against the fixture and the oracle.
"""
import os

import numpy as np
import pytest
import torch

from oracle import hno_oracle as orc

AUGMENT_CASES = orc.AUGMENT_CASES

from multimodal_3d_image_segmentation_b200.experiments.data_io import ImageTransform, draw_transform


def _load(golden_dir):
    return dict(np.load(os.path.join(golden_dir, 'augment.npz')))


def test_oracle_replays_reference_image_transform(golden_dir):
    g = _load(golden_dir)
    for name, spatial, kw, calls in AUGMENT_CASES:
        rng = np.random.default_rng(kw.get('seed'))
        okw = {k: v for k, v in kw.items() if k != 'seed'}
        nd = len(spatial)
        for c in range(calls):
            xo, yo, rec = orc.image_transform(g[f'{name}/x'], g[f'{name}/y'], rng, **okw)
            assert np.array_equal(xo, g[f'{name}/{c}/xo']) and np.array_equal(yo, g[f'{name}/{c}/yo']), (name, c)
            assert list(rec['flips']) == list(g[f'{name}/{c}/flips'])
            if f'{name}/{c}/matrix' in g:
                assert np.allclose(rec['matrix'], g[f'{name}/{c}/matrix'].reshape(nd, nd), rtol=0, atol=1e-12)
                assert np.allclose(rec['offset'], g[f'{name}/{c}/offset'], rtol=0, atol=1e-9)
            else:
                assert rec['matrix'] is None


def test_resampler_known_answers():
    """Hand-checkable cases of the nearest-neighbour resampler (the part SimpleITK owns in the reference)."""
    x = np.arange(2 * 3 * 4 * 5, dtype=np.float32).reshape(2, 3, 4, 5)
    assert np.array_equal(orc.affine_resample_nn(x, np.eye(3), np.zeros(3)), x)
    # output index p reads input p + (1, 0, -2) in (x, y, z): W shifts by one, D by minus two, cval outside
    y = orc.affine_resample_nn(x, np.eye(3), np.array([1.0, 0.0, -2.0]), cval=-7)
    assert np.array_equal(y[:, 2, :, :4], x[:, 0, :, 1:]) and np.all(y[:, :2] == -7) and np.all(y[:, :, :, 4] == -7)
    # round half up: a continuous index of exactly k + 0.5 reads voxel k + 1, and -0.5 is still inside (ITK IsInsideBuffer)
    y = orc.affine_resample_nn(x, np.eye(3), np.array([0.5, -0.5, 0.0]))
    assert np.array_equal(y[:, :, 0, :4], x[:, :, 0, 1:]) and np.all(y[:, :, :, 4] == 0)
    assert np.array_equal(y[:, :, 1:, :4], x[:, :, 1:, 1:])
    # 2-D image, axis swap
    im = np.arange(12, dtype=np.uint8).reshape(1, 3, 4)
    sw = orc.affine_resample_nn(im, np.array([[0.0, 1.0], [1.0, 0.0]]), np.zeros(2), cval=99)
    assert sw[0, 1, 2] == im[0, 2, 1] and sw[0, 0, 3] == 99


def test_host_draws_match_reference_matrices(golden_dir):
    """draw_transform (product host code) consumes the generator like the reference and composes the same matrices."""
    g = _load(golden_dir)
    for name, spatial, kw, calls in AUGMENT_CASES:
        tr = ImageTransform(**kw)
        nd = len(spatial)
        for c in range(calls):
            xform, flags = tr.draw(spatial)
            if f'{name}/{c}/matrix' in g:
                assert np.allclose(xform[:nd, :nd], g[f'{name}/{c}/matrix'].reshape(nd, nd), rtol=0, atol=1e-12), (name, c)
                assert np.allclose(xform[:nd, 3], g[f'{name}/{c}/offset'], rtol=0, atol=1e-9)
                if nd == 2:
                    assert np.array_equal(xform[2], [0, 0, 1, 0]) and np.all(xform[:2, 2] == 0)
            else:
                assert xform is None
            flips = [bool(flags >> (i + 3 - nd) & 1) for i in range(nd)]
            assert flips == [bool(v) for v in g[f'{name}/{c}/flips']], (name, c)


def test_no_cpu_path():
    tr = ImageTransform(flip=[True, True, True], seed=0)
    x = torch.zeros(1, 1, 2, 2, 2)
    with pytest.raises(RuntimeError):
        for _ in range(8):  # at least one of eight draws flips something
            tr.batch(x)


@pytest.mark.gpu
def test_device_transform_against_reference_fixture(cuda, golden_dir):
    g = _load(golden_dir)
    for name, spatial, kw, calls in AUGMENT_CASES:
        tr = ImageTransform(**kw)
        x = torch.from_numpy(g[f'{name}/x']).to(cuda)
        y = torch.from_numpy(g[f'{name}/y']).to(cuda)
        for c in range(calls):
            xo, yo = tr(x, y)
            assert xo.dtype == torch.float32 and yo.dtype == torch.uint8
            assert np.array_equal(xo.cpu().numpy(), g[f'{name}/{c}/xo']), (name, c)
            assert np.array_equal(yo.cpu().numpy(), g[f'{name}/{c}/yo']), (name, c)


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [torch.float32, torch.int16, torch.uint8, torch.int64])
def test_device_batch_against_oracle(cuda, dtype):
    """A batch: one parameter set per sample in generator order, every element type, odd sizes, non-zero cval."""
    kw = dict(rotation_range=[25, 10, 40], shift_range=[0.15, 0.1, 0.2], zoom_range=[0.75, 1.3], flip=[True, True, True],
              cval=3.0, augmentation_probability=0.85)
    B, C, spatial = 5, 3, (19, 23, 29)
    rng = np.random.default_rng(2)
    if dtype == torch.float32:
        x = rng.normal(size=(B, C) + spatial).astype(np.float32)
    else:
        x = rng.integers(0, 200, (B, C) + spatial).astype({torch.int16: np.int16, torch.uint8: np.uint8, torch.int64: np.int64}[dtype])
    lab = rng.integers(0, 4, (B, 1) + spatial).astype(np.uint8)
    tr = ImageTransform(seed=21, **kw)
    xo, yo = tr.batch(torch.from_numpy(x).to(cuda), torch.from_numpy(lab).to(cuda))
    assert xo.dtype == dtype and yo.dtype == torch.uint8
    orng = np.random.default_rng(21)
    for b in range(B):
        xr, yr, _ = orc.image_transform(x[b], lab[b], orng, **kw)
        assert np.array_equal(xo[b].cpu().numpy(), xr), b
        assert np.array_equal(yo[b].cpu().numpy(), yr), b


@pytest.mark.gpu
def test_device_transform_brats_grid_properties(cuda):
    """BASELINE-size volumes (4 x 240 x 240 x 155, batch 2): size-independent properties.  Flips are involutions, a pure
    integer shift is a slice copy, and resampling with the identity returns the input."""
    B, C, spatial = 2, 4, (240, 240, 155)
    g = torch.Generator(device='cuda').manual_seed(5)
    x = torch.randint(-300, 3000, (B, C) + spatial, device=cuda, dtype=torch.int16, generator=g)
    tr = ImageTransform()
    ident = np.hstack([np.eye(3), np.zeros((3, 1))])
    assert torch.equal(tr.batch(x, params=[(ident, 0)] * B), x)
    flipped = tr.batch(x, params=[(None, 0b101), (None, 0b010)])
    assert torch.equal(flipped[0], x[0].flip(1, 3)) and torch.equal(flipped[1], x[1].flip(2))
    assert torch.equal(tr.batch(flipped, params=[(None, 0b101), (None, 0b010)]), x)
    shift = ident.copy()
    shift[:, 3] = (3, -5, 7)  # (x, y, z): out[d, h, w] = in[d + 7, h - 5, w + 3]
    moved = tr.batch(x, params=[(shift, 0)] * B)
    assert torch.equal(moved[:, :, :233, 5:, :152], x[:, :, 7:, :235, 3:])
    assert int(moved[:, :, 233:].abs().max()) == 0 and int(moved[:, :, :, :5].abs().max()) == 0


def test_draw_transform_geometry_invariants():
    """Host logic: what the composed matrix must satisfy whatever the draws (reference dataset.py:128-165, 195-202)."""
    spatial = (30, 40, 50)
    centre = np.asarray(spatial[::-1], dtype=np.float64) / 2.0 + 0.5
    for seed in range(20):
        # rotation + zoom, no shift: a similarity about size / 2 + 0.5 -- the centre is a fixed point, A = zoom * R
        rng = np.random.default_rng(seed)
        xf, flags = draw_transform(rng, spatial, rotation_range=[30, 20, 10], zoom_range=[0.7, 1.4])
        A, t = xf[:, :3], xf[:, 3]
        assert flags == 0 and np.allclose(A @ centre + t, centre, atol=1e-9)
        zoom = np.cbrt(np.linalg.det(A))
        assert 0.7 <= zoom <= 1.4 and np.allclose(A @ A.T, zoom * zoom * np.eye(3), atol=1e-12)
        # shift only: identity matrix, offset = the drawn fractions of the size in (x, y, z) order
        rng = np.random.default_rng(seed)
        xf, _ = draw_transform(rng, spatial, shift_range=[0.1, 0.2, 0.3])
        assert np.array_equal(xf[:, :3], np.eye(3))
        frac = xf[:, 3] / np.asarray(spatial[::-1], dtype=np.float64)
        assert np.all(np.abs(frac) <= np.array([0.3, 0.2, 0.1]) + 1e-12)
    # the probability gate: nothing else is drawn when it says no
    rng = np.random.default_rng(0)
    none = [draw_transform(rng, spatial, rotation_range=[30, 30, 30], flip=[True] * 3, augmentation_probability=0.0)
            for _ in range(5)]
    assert all(xf is None and fl == 0 for xf, fl in none)
    ref = np.random.default_rng(0)
    for _ in range(5):
        ref.binomial(1, 0.0)
    assert rng.random() == ref.random()  # the generator advanced by exactly the five gates


def test_resampler_shift_round_trip_and_zoom():
    rng = np.random.default_rng(4)
    x = rng.integers(1, 100, (2, 9, 10, 11)).astype(np.int16)
    eye = np.eye(3)
    there = orc.affine_resample_nn(x, eye, np.array([2.0, -1.0, 3.0]))
    back = orc.affine_resample_nn(there, eye, np.array([-2.0, 1.0, -3.0]))
    # voxels that never left the volume come back unchanged, the rest were filled with cval = 0
    assert np.array_equal(back[:, 3:, :9, 2:], x[:, 3:, :9, 2:])
    assert np.all(back[:, :3] == 0) and np.all(back[:, :, 9:] == 0) and np.all(back[:, :, :, :2] == 0)
    # zoom by exactly 2 about the origin: output p reads input 2 p (nearest neighbour = exact sub-sampling)
    z = orc.affine_resample_nn(x, 2.0 * eye, np.zeros(3), cval=-1)
    assert np.array_equal(z[:, :5, :5, :6], x[:, ::2, ::2, ::2]) and np.all(z[:, 5:] == -1)
