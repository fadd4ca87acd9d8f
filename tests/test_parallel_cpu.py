"""Host-side data-parallel logic on CPU: world_size 2, gloo backend (the GPU box runs the same code over NCCL)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _make_module():
    torch.manual_seed(7)
    return torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from multimodal_3d_image_segmentation_b200 import parallel
    model = _make_module()
    flat = parallel.FlatParameters(model)
    # parameters are views of ONE buffer, gradients of another
    assert all(p.data_ptr() >= flat.data.data_ptr() for p in model.parameters())
    assert flat.numel == sum(p.numel() for p in model.parameters())
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    parallel.attach_gradient_allreduce(opt, flat)
    g = torch.Generator().manual_seed(100 + rank)  # this rank's shard of the global batch
    x = torch.randn(4, 5, generator=g)
    loss = model(x).pow(2).mean()
    flat.grad.zero_()
    loss.backward()  # accumulates into the flat gradient views
    opt.step()       # pre-hook: ONE all-reduce + 1/world
    torch.save({'params': flat.data.clone(), 'grad': flat.grad.clone()}, os.path.join(out_dir, f'r{rank}.pt'))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_rank_on_the_concatenated_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f'r{r}.pt')) for r in range(world)]
    # replicas stay identical
    assert torch.equal(res[0]['params'], res[1]['params'])
    assert torch.equal(res[0]['grad'], res[1]['grad'])
    # and equal the single-process step on the whole batch (mean of per-shard mean losses)
    model = _make_module()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    xs = [torch.randn(4, 5, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    loss = sum(model(x).pow(2).mean() for x in xs) / world
    loss.backward()
    opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert torch.allclose(res[0]['params'], ref, rtol=1e-6, atol=1e-7)


def test_allreduce_is_identity_without_process_group():
    from multimodal_3d_image_segmentation_b200 import parallel
    g = torch.arange(6, dtype=torch.float32)
    assert torch.equal(parallel.allreduce_mean_(g.clone()), g)
