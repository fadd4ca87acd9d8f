"""One launch of each kernel added in the last widening of round 2, at BASELINE sizes (for ncu; also prints CUDA-event times):
k_affine_resample_nn (fp32 images, uint8 labels), k_head_direct_fwd / _bwd (use_resize=False head), k_stem_dx (image gradient)."""
import json
import sys

import torch

sys.path.insert(0, '.')
from multimodal_3d_image_segmentation_b200 import ops  # noqa: E402
from multimodal_3d_image_segmentation_b200.experiments.data_io import ImageTransform  # noqa: E402
from multimodal_3d_image_segmentation_b200.plan import plane_pitch  # noqa: E402

dev = torch.device('cuda:0')
B, C, sp = 2, 4, (240, 240, 155)
N = sp[0] * sp[1] * sp[2]
x = torch.randn(B, C, *sp, device=dev)
y = torch.randint(0, 4, (B, 1) + sp, device=dev, dtype=torch.uint8)
tr = ImageTransform(rotation_range=[30, 30, 30], shift_range=[0.2, 0.2, 0.2], zoom_range=[0.8, 1.2], seed=1)
params = [tr.draw(sp) for _ in range(B)]
logits = torch.randn(B, C, sp[0], sp[1] * sp[2], device=dev)
probs = ops.head_direct_forward(logits, sp, 1)
dprobs = torch.randn_like(probs)
D, H, W = ops.stem_out_shape(sp)
pitch = plane_pitch(H, W)
dpre = torch.randn(B, 24, D, pitch, device=dev)
wstem = torch.randn(24, C, 2, 2, 2, device=dev)
xz = torch.randn(B, C, 155, 240, 240, device=dev)  # SimpleITK (z, y, x) order


def timed(fn, n=10):
    if len(sys.argv) > 1 and sys.argv[1] == 'once':  # under ncu: exactly one launch of everything
        fn()
        torch.cuda.synchronize()
        return 1.0
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows = []
for name, fn, nbytes in (
        ('affine_resample_nn fp32 images + u8 labels', lambda: tr.batch(x, y, params=params), 2 * (x.numel() * 4 + y.numel())),
        ('head_direct_forward (softmax)', lambda: ops.head_direct_forward(logits, sp, 1), 2 * B * C * N * 4),
        ('head_direct_backward', lambda: ops.head_direct_backward(dprobs, probs, sp[1] * sp[2], 1), 3 * B * C * N * 4),
        ('stem_backward_input', lambda: ops.stem_backward_input(dpre, wstem, sp, pitch), dpre.numel() * 4 + B * C * N * 4),
        ('permute_spatial (155,240,240)->(240,240,155) fp32', lambda: ops.permute_spatial(xz, (1, 2, 0)), 2 * xz.numel() * 4)):
    ms = timed(fn)
    rows.append({'kernel': name, 'ms': round(ms, 4), 'alg_bytes': nbytes, 'GB_per_s': round(nbytes / ms / 1e6, 1)})
    print(json.dumps(rows[-1]))
