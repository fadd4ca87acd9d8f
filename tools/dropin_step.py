"""Drop-in mode: the step body of the reference's experiments/train_test.py:146-175, line for line, on this package's
modules -- pinned host batch -> device, to_categorical, model(x) -> probabilities, loss_fn(y_pred, y_onehot),
loss.item(), zero_grad, backward, torch.optim.Adamax.step, CosineAnnealingWarmRestarts.step -- timed with CUDA events
next to the fused Trainer.step on the same batch.  This is what `experiments/run.py config_hnoseg_xs.ini` gets after the
package swap of INTEGRATION.md section 1 with no other change.
Usage: python tools/dropin_step.py [--steps 10] [--loss PCCLoss] [out.json]"""
import argparse
import json
import sys

import torch

sys.path.insert(0, '.')
from multimodal_3d_image_segmentation_b200 import nets  # noqa: E402
from multimodal_3d_image_segmentation_b200.experiments import to_categorical  # noqa: E402
from multimodal_3d_image_segmentation_b200.parallel import Trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=2)
    ap.add_argument('--loss', default='PCCLoss')  # config_hnoseg_xs.ini:61-62
    ap.add_argument('out', nargs='?')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    model = nets.HNOSegXS(4, 4, 24, [3] * 8, (10, 14, 14), device=dev)
    loss_fn = getattr(nets.custom_losses, a.loss)()
    optimizer = torch.optim.Adamax(model.parameters(), lr=5e-3)
    scheduler = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(optimizer, T_0=1000, eta_min=1e-3)
    shape = (240, 240, 155)
    g = torch.Generator().manual_seed(1234)
    xh = torch.randn(a.batch, 4, *shape, generator=g).pin_memory()
    yh = torch.randint(0, 4, (a.batch, 1, *shape), generator=g, dtype=torch.uint8).pin_memory()
    num_labels = model.out_channels

    def step():  # train_test.py:146-175
        x = xh.to(dev, non_blocking=True)
        y = yh.to(dev, non_blocking=True)
        y = to_categorical(y, num_labels)
        y_pred = model(x)
        loss = loss_fn(y_pred, y)
        value = loss.item()
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        scheduler.step()
        return value

    def timed(fn):
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            last = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps, last

    ms_dropin, loss_dropin = timed(step)
    del optimizer, scheduler
    torch.cuda.empty_cache()
    trainer = Trainer(model, loss_name=a.loss)
    xd, yd = xh.to(dev), yh.to(dev)

    def fused():
        xd.copy_(xh, non_blocking=True)
        yd.copy_(yh, non_blocking=True)
        return trainer.step(xd, yd).item()

    ms_fused, loss_fused = timed(fused)
    res = {'config': f'HNOSegXS(4,4,24,[3]*8,(10,14,14)), batch {a.batch} x 4x240x240x155, {a.loss}, host batch copied every step',
           'dropin': {'ms_per_step': round(ms_dropin, 3), 'volumes_per_s': round(a.batch / ms_dropin * 1e3, 1), 'loss': loss_dropin,
                      'what': 'train_test.py:146-175 body: to_categorical, model(x) -> probabilities, loss_fn, loss.item(), '
                              'autograd backward, torch.optim.Adamax, scheduler'},
           'fused': {'ms_per_step': round(ms_fused, 3), 'volumes_per_s': round(a.batch / ms_fused * 1e3, 1), 'loss': loss_fused,
                     'what': 'parallel.Trainer.step (uint8 labels, fused head + loss, CUDA graph, fused Adamax), serial copies'},
           'steps': a.steps, 'warmup': a.warmup, 'peak_mem_GiB': round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
    print(json.dumps(res, indent=1))
    if a.out:
        json.dump(res, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
