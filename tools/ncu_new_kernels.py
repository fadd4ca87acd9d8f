"""Launches each loss / input-side kernel added at the end of round 1 a few times at the BASELINE volume size, for
  ncu --set full --clock-control none -k regex:"k_ce_|k_loss_moments|k_to_categorical|k_norm_" -s 16 -c 8 -o out python tools/ncu_new_kernels.py
(two warm passes = 16 matching launches are skipped, the third pass is captured)."""
import sys

import torch

sys.path.insert(0, '.')
from multimodal_3d_image_segmentation_b200 import nets, ops  # noqa: E402
from multimodal_3d_image_segmentation_b200.experiments import normalize_modalities, to_categorical  # noqa: E402

dev = torch.device('cuda:0')
B, C, shape = 2, 4, (240, 240, 155)
p = torch.softmax(torch.randn(B, C, *shape, device=dev), 1)
lab = torch.randint(0, C, (B,) + shape, device=dev, dtype=torch.uint8)
vol = torch.rand(4, 155, 240, 240, device=dev) * 1000
vol *= (torch.rand(1, 155, 240, 240, device=dev) < 0.4)
dice = nets.custom_losses.DiceLoss()
for _ in range(3):
    onehot = to_categorical(lab[:, None], C, validate=False)   # k_to_categorical_u8<4>
    ops.ce_loss_forward(p, labels=lab)                         # k_ce_fwd<4,1,4> (+ k_ce_finalize)
    ops.ce_loss_backward(p, labels=lab)                        # k_ce_bwd<4,1,4>
    dice(p, onehot)                                            # k_loss_moments<4> (+ k_loss_finalize)
    normalize_modalities(vol, mask_val=0)                      # k_norm_moments<4>, k_norm_finalize, k_norm_apply<4>
    torch.cuda.synchronize()
