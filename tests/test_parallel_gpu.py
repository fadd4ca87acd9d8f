"""Numerical parity of the data-parallel step over NCCL: N ranks on shards == one rank on the whole batch.

Needs >= 2 GPUs (skipped otherwise): `gpurun --gpus 2 -- python -m pytest tests/test_parallel_gpu.py -m gpu`.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

CFG = dict(in_channels=4, out_channels=4, filters=24, num_transform_blocks=[3] * 8, num_modes=(10, 14, 14))
SHAPE = (64, 56, 48)
PER_RANK = 2
STEPS = 3


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _batch(rank, step):
    g = torch.Generator().manual_seed(1000 + 17 * rank + step)
    x = torch.randn(PER_RANK, 4, *SHAPE, generator=g)
    lab = torch.randint(0, 4, (PER_RANK, 1, *SHAPE), generator=g).to(torch.uint8)
    return x, lab


def _worker(rank, world, port, out_dir, loss_name):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from multimodal_3d_image_segmentation_b200 import nets, parallel
    from oracle import hno_oracle as orc
    model = nets.HNOSegXS(**CFG, device=dev)
    model.load_state_dict(orc.init_state_dict(4, 4, 24, [3] * 8, (10, 14, 14), seed=0))
    tr = parallel.Trainer(model, loss_name, lr=5e-3)
    losses, grads, before, after = [], [], [], []
    for step in range(STEPS):
        x, lab = _batch(rank, step)
        before.append(tr.flat.data.cpu().clone())
        losses.append(float(tr.step(x.to(dev), lab.to(dev))))
        grads.append(tr.flat.grad.cpu().clone())  # after the all-reduce: the averaged gradient the optimizer used
        after.append(tr.flat.data.cpu().clone())
    torch.save({'losses': losses, 'grads': grads, 'before': before, 'after': after}, os.path.join(out_dir, f'r{rank}.pt'))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('loss_name', ['DiceLoss', 'PCCLoss'])
def test_nccl_two_rank_step_equals_single_rank_on_the_concatenated_batch(cuda, tmp_path, loss_name):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), loss_name), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f'r{r}.pt')) for r in range(world)]
    for step in range(STEPS):  # replicas stay bit-identical: same averaged gradient, same parameters, every step
        assert torch.equal(res[0]['grads'][step], res[1]['grads'][step])
        assert torch.equal(res[0]['after'][step], res[1]['after'][step])
    from multimodal_3d_image_segmentation_b200 import nets, parallel
    from oracle import hno_oracle as orc
    model = nets.HNOSegXS(**CFG, device=cuda)
    model.load_state_dict(orc.init_state_dict(4, 4, 24, [3] * 8, (10, 14, 14), seed=0))
    tr = parallel.Trainer(model, loss_name, lr=5e-3)
    # (a) At the parameters the two ranks held before each step, ONE rank on the concatenated batch computes the same loss and
    # the same gradient as the all-reduced mean of the two shards, up to fp32 summation order (the rank partial sums are combined
    # by NCCL instead of inside one kernel).  The global loss is the mean of the rank losses (means over (sample, label),
    # nets/custom_losses.py:70,111).  Comparing free-running trajectories instead is ill-posed: Adamax's first update is
    # lr * sign(g) for every element, so an element whose gradient is at round-off level moves 2 lr apart (measured 1.2e-3
    # relative after three steps, later gradients 0.14 apart).
    for step in range(STEPS):
        xs, labs = zip(*[_batch(r, step) for r in range(world)])
        tr.flat.data.copy_(res[0]['before'][step])
        loss = float(tr.loss_and_grad(torch.cat(xs).to(cuda), torch.cat(labs).to(cuda)))
        dl = abs(loss - (res[0]['losses'][step] + res[1]['losses'][step]) / 2)
        g = tr.flat.grad.cpu()
        dg = ((g - res[0]['grads'][step]).norm() / g.norm()).item()
        print(f'{loss_name}: step {step + 1}: loss {loss:.7f} (2-rank mean differs by {dl:.1e}), gradient rel-L2 {dg:.2e}')
        assert dl < 2e-6, (step, dl)
        assert dg < 1e-5, (step, dg)
    # (b) Given the all-reduced gradients, the fused Adamax reproduces the ranks' parameters bit for bit.
    model2 = nets.HNOSegXS(**CFG, device=cuda)
    model2.load_state_dict(orc.init_state_dict(4, 4, 24, [3] * 8, (10, 14, 14), seed=0))
    tr2 = parallel.Trainer(model2, loss_name, lr=5e-3)
    assert torch.equal(tr2.flat.data.cpu(), res[0]['before'][0])
    for step in range(STEPS):
        tr2.flat.grad.copy_(res[0]['grads'][step])
        tr2.optimizer.step()
        assert torch.equal(tr2.flat.data.cpu(), res[0]['after'][step]), step
