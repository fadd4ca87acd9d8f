"""Times the stand-alone loss and input-side kernels at the BASELINE volume size (2 x 4 x 240 x 240 x 155) with CUDA events.
Algorithmic bytes: moments pass P + P (one-hot floats), CE forward P + T, CE backward P + T + P, T = one-hot floats (P)
or uint8 labels (P / 16).  Usage: python tools/time_losses.py [out.json]"""
import json
import sys

import torch

sys.path.insert(0, '.')
from multimodal_3d_image_segmentation_b200 import ops  # noqa: E402
from multimodal_3d_image_segmentation_b200 import nets  # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dev = torch.device('cuda:0')
    B, C, shape = 2, 4, (240, 240, 155)
    N = shape[0] * shape[1] * shape[2]
    p = torch.softmax(torch.randn(B, C, *shape, device=dev), 1)
    lab = torch.randint(0, C, (B,) + shape, device=dev, dtype=torch.uint8)
    onehot = torch.zeros_like(p).scatter_(1, lab[:, None].long(), 1.0)
    P = B * C * N * 4
    L = B * N
    rows = {}

    def add(name, ms, nbytes):
        rows[name] = {'ms': round(ms, 4), 'algorithmic_MB': round(nbytes / 1e6, 1), 'GB_per_s': round(nbytes / ms / 1e6, 1)}

    add('ce_fwd_labels', timed(lambda: ops.ce_loss_forward(p, labels=lab)), P + L)
    add('ce_fwd_onehot', timed(lambda: ops.ce_loss_forward(p, y_true=onehot)), 2 * P)
    add('ce_bwd_labels', timed(lambda: ops.ce_loss_backward(p, labels=lab)), 2 * P + L)
    add('ce_bwd_onehot', timed(lambda: ops.ce_loss_backward(p, y_true=onehot)), 3 * P)
    for name in ('DiceLoss', 'PCCLoss', 'ExpDiceLoss'):
        fn = getattr(nets.custom_losses, name)()
        add(name + '_fwd_onehot', timed(lambda: fn(p, onehot)), 2 * P)
    # input side: one-hot from uint8 labels (reads L, writes P), modality normalisation of one
    # 4 x 155 x 240 x 240 sample with the background masked (2 reads + 1 write of the 142.8 MB volume)
    from multimodal_3d_image_segmentation_b200.experiments import normalize_modalities, to_categorical
    lab5 = lab[:, None]
    add('to_categorical_u8', timed(lambda: to_categorical(lab5, C, validate=False)), P + L)
    vol = torch.rand(4, 155, 240, 240, device=dev) * 1000
    vol *= (torch.rand(1, 155, 240, 240, device=dev) < 0.4)
    add('normalize_modalities_mask0', timed(lambda: normalize_modalities(vol, mask_val=0)), 3 * vol.numel() * 4)
    print(json.dumps(rows, indent=1))
    if len(sys.argv) > 1:
        json.dump(rows, open(sys.argv[1], 'w'), indent=1)


if __name__ == '__main__':
    main()
