"""Two training steps of the BASELINE config (batch 2) for ncu: the first is warm-up, profile the second."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_3d_image_segmentation_b200 import nets, parallel  # noqa: E402
from oracle import hno_oracle as orc  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device('cuda:0')
cfg = dict(in_channels=4, out_channels=4, filters=24, num_transform_blocks=[3] * 8, num_modes=(10, 14, 14))
model = nets.HNOSegXS(**cfg, device=dev)
model.load_state_dict(orc.init_state_dict(4, 4, 24, [3] * 8, (10, 14, 14), seed=0))
trainer = parallel.Trainer(model, "DiceLoss", use_graph=False)
g = torch.Generator().manual_seed(1234)
x = torch.randn(batch, 4, 240, 240, 155, generator=g).to(dev)
lab = torch.randint(0, 4, (batch, 1, 240, 240, 155), generator=g).to(torch.uint8).to(dev)
for i in range(steps):
    parallel.launches(reset=True)
    loss = trainer.step(x, lab)
    torch.cuda.synchronize()
    print(f'step {i}: loss {float(loss):.6f}, launches {parallel.launches()}', flush=True)
