"""Runs the fused head + loss forward / backward at the BASELINE shapes (batch 2) a few times, for ncu captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_3d_image_segmentation_b200 import ops  # noqa: E402
from multimodal_3d_image_segmentation_b200.plan import get_interp_tables, plane_pitch  # noqa: E402

dev = torch.device('cuda:0')
VOLUME, batch, C = (240, 240, 155), 2, 4
D, H, W = ops.stem_out_shape(VOLUME)
P = plane_pitch(H, W)
tables = get_interp_tables((D, H, W), VOLUME, dev)
g = torch.Generator(device=dev).manual_seed(3)
ll = torch.randn(batch, C, D, P, device=dev, generator=g)
lab = torch.randint(0, 4, (batch, 1, *VOLUME), device=dev, generator=g).to(torch.uint8)
for name in ('fwd', 'bwd'):
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if name == 'fwd':
            loss, coef = ops.head_loss_forward(ll, lab, tables, P, 0)
        else:
            dll = ops.head_loss_backward(ll, lab, coef, None, tables, P)
        e1.record()
        torch.cuda.synchronize()
    print(f'{name}: {e0.elapsed_time(e1):.4f} ms', flush=True)
