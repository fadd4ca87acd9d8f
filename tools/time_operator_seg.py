"""Training-step time of the NeuralOperatorSeg family at the shipped configs (experiments/config_files/config_hnoseg.ini,
config_fnoseg.ini, config_fno.ini) on one 4 x 240 x 240 x 155 volume: model(x) -> PCCLoss -> backward through the drop-in
modules (one autograd node per op; these models have no fused engine yet), CUDA events, device-resident batch.
Usage: python tools/time_operator_seg.py [out.json]"""
import json
import sys

import torch

sys.path.insert(0, '.')
from multimodal_3d_image_segmentation_b200 import nets  # noqa: E402
from multimodal_3d_image_segmentation_b200.experiments import to_categorical  # noqa: E402

CONFIGS = {
    'config_hnoseg.ini (HNOSeg, Hartley, shared, modes (10,14,14), 24 blocks)':
        dict(num_transform_blocks=24, num_modes=(10, 14, 14), transform_type='Hartley'),
    'config_fnoseg.ini (FNOSeg, Fourier, shared, modes (10,14,14), 24 blocks)':
        dict(num_transform_blocks=24, num_modes=(10, 14, 14), transform_type='Fourier'),
    'config_fno.ini (FNO, Fourier, individual, modes (4,6,6), 24 blocks, no block skip)':
        dict(num_transform_blocks=24, num_modes=(4, 6, 6), transform_type='Fourier', weights_type='individual',
             use_bias_conv_branch=True, use_block_skip=False),
}


def main():
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, 4, 240, 240, 155, generator=g).to(dev)
    y = to_categorical(torch.randint(0, 4, (1, 1, 240, 240, 155), generator=g, dtype=torch.uint8).to(dev), 4)
    loss_fn = nets.custom_losses.PCCLoss()
    res = {}
    for name, kw in CONFIGS.items():
        torch.manual_seed(0)
        model = nets.NeuralOperatorSeg(4, 4, 24, device=dev, **kw)
        nparam = sum(p.numel() for p in model.parameters())

        def step():
            model.zero_grad(set_to_none=True)
            loss = loss_fn(model(x), y)
            loss.backward()
            return loss

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 4
        e0.record()
        for _ in range(n):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        with torch.no_grad():
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(n):
                model(x)
            f1.record()
            torch.cuda.synchronize()
        res[name] = {'params': nparam, 'train_ms_per_volume': round(ms, 2), 'train_volumes_per_s': round(1e3 / ms, 1),
                     'forward_ms_per_volume': round(f0.elapsed_time(f1) / n, 2), 'loss': float(loss),
                     'peak_mem_GiB': round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
        print(name, res[name], flush=True)
        del model
        torch.cuda.empty_cache()
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], 'w'), indent=1)


if __name__ == '__main__':
    main()
