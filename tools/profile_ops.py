"""Runs selected hot-path entry points at the BASELINE shapes (batch 2) a few times, for ncu captures:
    python tools/profile_ops.py pw48f dhtf dhts dhta pw48b mhaf mhab [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_3d_image_segmentation_b200 import ops  # noqa: E402
from multimodal_3d_image_segmentation_b200.plan import get_crop_plan, plane_pitch  # noqa: E402

names = [a for a in sys.argv[1:] if not a.isdigit()]
reps = next((int(a) for a in sys.argv[1:] if a.isdigit()), 2)
dev = torch.device('cuda:0')
VOLUME, MODES, F, batch = (240, 240, 155), (10, 14, 14), 24, 2
D, H, W = ops.stem_out_shape(VOLUME)
P = plane_pitch(H, W)
plan = get_crop_plan((D, H, W), MODES, dev)
g = torch.Generator(device=dev).manual_seed(3)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
a = [rnd(batch, F, D, P) for _ in range(4)]
acc = [rnd(batch, F, D, P) for _ in range(2)]
z = rnd(batch, F, *plan.modes_shape)
w48, w24, b24 = rnd(F, 2 * F) * 0.1, rnd(F, F) * 0.1, rnd(F) * 0.01
hw = (P, H * W)
xin = rnd(batch, 4, *VOLUME)
zm = rnd(batch, 12, 20, 28, 28)
wm = [rnd(4, 12, 12) * 0.1 for _ in range(3)] + [rnd(12, 48) * 0.1]
_mha = {}


def mha_fwd():
    _mha['y'], _mha['S'] = ops.hartley_attention_forward(zm, None, None, *wm, patch=(2, 2, 2), activation=1)


def mha_bwd():
    if 'S' not in _mha:
        mha_fwd()
    ops.hartley_attention_backward(_mha['y'], _mha['S'])

win = rnd(F, 4, 2, 2, 2) * 0.1
wch = [w24, w24 * 0.5, w24 * 0.25]
_ch = {}


def chain_fwd():
    _ch['u'], _ch['z'] = ops.dht3_chain_forward(a[0], plan, wch, 1.0 / plan.n_voxels, epilogue=2, save=True)


def chain_bwd():
    if 'z' not in _ch:
        chain_fwd()
    ops.dht3_chain_backward(a[2], plan, _ch['z'], wch, 1.0 / plan.n_voxels, a[3], epilogue=1)

table = {
    'stemf': lambda: ops.stem_forward(xin, win, b24, P),
    'stemb': lambda: ops.stem_backward(a[0], xin, F, P),
    'pw48f': lambda: ops.pwconv_forward(a[0], a[1], w48, b24, 1, False),
    'pw24f': lambda: ops.pwconv_forward(a[0], None, w24, b24, 1, False),
    'dhtf': lambda: ops.dht3_forward(a[0], plan, 1.0),
    'dhts': lambda: ops.dht3_adjoint(z, plan, 1.0, epilogue=2, out=a[1]),
    'dhtp': lambda: ops.dht3_adjoint(z, plan, 1.0, epilogue=0, out=a[1]),
    'dhta': lambda: ops.dht3_adjoint(z, plan, 1.0, epilogue=1, out=a[1]),
    'pw48b': lambda: ops.pwconv_backward(a[0], a[1], a[2], a[3], w48, 1, False, hw=hw, in1_is_selu=True),
    'pw48ba': lambda: ops.pwconv_backward(a[0], a[1], a[2], a[3], w48, 1, False, hw=hw, din1=acc[0], din2=acc[1]),
    'chainf': chain_fwd,
    'chainb': chain_bwd,
    'mhaf': mha_fwd,
    'mhab': mha_bwd,
    'pw24b': lambda: ops.pwconv_backward(a[0], a[1], a[2], None, w24, 1, False, hw=hw, in1_is_selu=True),
}
for n in names:
    for _ in range(reps):
        table[n]()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        table[n]()
    ev[1].record()
    torch.cuda.synchronize()
    print(f'{n}: {ev[0].elapsed_time(ev[1]) / 5:.4f} ms', flush=True)
