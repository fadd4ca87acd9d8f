"""Times the on-device augmentation (hno_affine_resample_nn) at the BASELINE batch: 2 x 4 x 240 x 240 x 155 raw int16
modalities + uint8 labels, parameters drawn like config_hnoseg_xs.ini's [augmentation].  CUDA events, 20 launches."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from multimodal_3d_image_segmentation_b200.experiments.data_io import ImageTransform  # noqa: E402

dev = torch.device('cuda:0')
B, C, sp = 2, 4, (240, 240, 155)
x = torch.randint(0, 3000, (B, C) + sp, device=dev, dtype=torch.int16)
y = torch.randint(0, 4, (B, 1) + sp, device=dev, dtype=torch.uint8)
tr = ImageTransform(rotation_range=[30, 30, 30], shift_range=[0.2, 0.2, 0.2], zoom_range=[0.8, 1.2], seed=1)
params = [tr.draw(sp) for _ in range(B)]
for _ in range(3):
    tr.batch(x, y, params=params)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    tr.batch(x, y, params=params)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
nbytes = 2 * (x.numel() * 2 + y.numel())  # one read + one write of images and labels
print(json.dumps({'op': 'augment_batch', 'ms': round(ms, 4), 'GB_per_s': round(nbytes / ms / 1e6, 1), 'bytes': nbytes,
                  'note': 'includes the two small H2D parameter copies per call'}))
xf = x.float()
for _ in range(3):
    tr.batch(xf, params=params)
torch.cuda.synchronize()
e0.record()
for _ in range(n):
    tr.batch(xf, params=params)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({'op': 'augment_images_fp32', 'ms': round(ms, 4), 'GB_per_s': round(2 * xf.numel() * 4 / ms / 1e6, 1)}))
