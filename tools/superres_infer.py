"""BASELINE config 4: HNOSeg-XS zero-shot super-resolution inference at the 2x grid (4 x 480 x 480 x 310) on one B200:
forward time per volume with the input resident in HBM (CUDA events) and end to end from pinned host memory with the
uint8 label map read back (the reference's testing() loop does the argmax on the host, train_test.py:398-408)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_3d_image_segmentation_b200 import nets  # noqa: E402
from oracle import hno_oracle as orc  # noqa: E402

dev = torch.device('cuda:0')
modes = (10, 14, 14)
model = nets.HNOSegXS(4, 4, 24, [3] * 8, modes, device=dev).eval()
model.load_state_dict(orc.init_state_dict(4, 4, 24, [3] * 8, modes, seed=0))
for shape in ((240, 240, 155), (480, 480, 310)):
    xh = torch.randn(1, 4, *shape, generator=torch.Generator().manual_seed(9)).pin_memory()
    x = xh.to(dev)
    with torch.no_grad():
        for _ in range(2):
            model(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            probs = model(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        e0.record()
        for _ in range(3):
            lab = model(xh.to(dev, non_blocking=True)).argmax(1).to(torch.uint8).cpu()
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 3
    print(f'{shape}: forward {ms:.2f} ms/volume ({1e3 / ms:.1f} volumes/s) resident; {ms2:.2f} ms/volume '
          f'({1e3 / ms2:.1f} volumes/s) host -> labels on host; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB',
          flush=True)
