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
    # two steps exactly as the reference loop writes them (experiments/train_test.py:164-171): zero_grad() defaults to
    # set_to_none=True, which detaches p.grad from the flat buffer; the hook has to re-pack before it reduces
    for it in range(2):
        loss = model(x + it).pow(2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()       # pre-hook: re-pack + ONE all-reduce + 1/world
        assert all(p.grad.data_ptr() == gv.data_ptr() for p, gv in zip(flat.params, flat.grad_views))
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
    for it in range(2):
        loss = sum(model(x + it).pow(2).mean() for x in xs) / world
        opt.zero_grad()
        loss.backward()
        opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert torch.allclose(res[0]['params'], ref, rtol=1e-6, atol=1e-7)


def test_allreduce_is_identity_without_process_group():
    from multimodal_3d_image_segmentation_b200 import parallel
    g = torch.arange(6, dtype=torch.float32)
    assert torch.equal(parallel.allreduce_mean_(g.clone()), g)


def test_repack_after_zero_grad_set_to_none():
    """optimizer.zero_grad() drops p.grad; backward then allocates gradients outside the flat buffer."""
    from multimodal_3d_image_segmentation_b200 import parallel
    model = _make_module()
    flat = parallel.FlatParameters(model)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    opt.zero_grad()
    assert all(p.grad is None for p in model.parameters())
    model(torch.ones(2, 5)).sum().backward()
    assert flat.grad.abs().sum() == 0  # the situation the hook has to repair
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    flat.repack_grads_()
    assert torch.equal(flat.grad, want)
    assert all(p.grad.data_ptr() == gv.data_ptr() for p, gv in zip(flat.params, flat.grad_views))
    # a parameter without gradient contributes zeros
    opt.zero_grad()
    model[0](torch.ones(2, 5)).sum().backward()
    flat.repack_grads_()
    assert flat.grad_views[2].abs().sum() == 0 and flat.grad_views[0].abs().sum() > 0


def test_fused_adamax_state_dict_is_a_snapshot_and_converts_to_torch_layout():
    from multimodal_3d_image_segmentation_b200 import parallel
    model = _make_module()
    flat = parallel.FlatParameters(model)
    opt = parallel.FusedAdamax(flat, lr=1e-2)
    opt.exp_avg.fill_(0.5)
    opt.exp_inf.fill_(0.25)
    opt.step_count = 7
    sd = opt.state_dict()
    opt.exp_avg.add_(1.0)
    assert torch.all(sd['exp_avg'] == 0.5)  # cloned, not the live tensor
    tsd = opt.torch_state_dict()
    ref = torch.optim.Adamax(_make_module().parameters(), lr=5e-3)
    ref.load_state_dict(tsd)  # the reference's optimizer accepts it (train_test.py:276-286)
    assert ref.state_dict()['param_groups'][0]['lr'] == 1e-2
    st = ref.state_dict()['state']
    assert sorted(st) == [0, 1, 2, 3] and float(st[1]['step']) == 7 and st[1]['exp_avg'].shape == (3,)
    opt2 = parallel.FusedAdamax(parallel.FlatParameters(_make_module()), lr=5e-3)
    opt2.load_state_dict(ref.state_dict())  # and back
    assert opt2.step_count == 7 and opt2.lr == 1e-2
    assert torch.all(opt2.exp_avg == 1.5) and torch.all(opt2.exp_inf == 0.25)


def test_cosine_warm_restarts_matches_torch_scheduler():
    """Host-side schedule for FusedAdamax == torch's CosineAnnealingWarmRestarts stepped per batch (run.py:96-103)."""
    from multimodal_3d_image_segmentation_b200 import parallel
    for T_0, T_mult, eta_min in ((7, 1, 0.0), (5, 2, 1e-4), (3, 3, 0.0)):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adamax([p], lr=5e-3)
        sch = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=T_0, T_mult=T_mult, eta_min=eta_min)
        for step in range(60):
            want = opt.param_groups[0]['lr']
            got = parallel.cosine_warm_restarts_lr(step, 5e-3, T_0, T_mult, eta_min)
            assert abs(got - want) < 1e-12, (T_0, T_mult, step, got, want)
            opt.step()
            sch.step()


def test_axis_permutation_policy(monkeypatch):
    """Trainer._axis_perm (host logic): large volumes whose last axis is not the shortest run with the shortest axis last;
    the BASELINE-literal order, small volumes, per-mode weights and use_resize=False are left alone."""
    import types
    from multimodal_3d_image_segmentation_b200.parallel import Trainer
    monkeypatch.delenv('HNO_AXIS_PERM', raising=False)

    def policy(spatial, **kw):
        model = types.SimpleNamespace(use_resize=kw.get('use_resize', True), weights_type=kw.get('weights_type', 'shared'))
        return Trainer._axis_perm(types.SimpleNamespace(_direct=kw.get('direct', True), model=model), spatial)

    assert policy((240, 240, 155)) is None            # BASELINE-literal order: already shortest last
    assert policy((155, 240, 240)) == (1, 2, 0)       # SimpleITK (z, y, x) order of a BraTS volume
    assert policy((240, 155, 240)) == (0, 2, 1)
    assert policy((13, 16, 18)) is None               # small: the transposes cost more than they save
    assert policy((155, 240, 240), weights_type='individual') is None
    assert policy((155, 240, 240), use_resize=False) is None
    assert policy((155, 240, 240), direct=False) is None
    monkeypatch.setenv('HNO_AXIS_PERM', '0')
    assert policy((155, 240, 240)) is None
    monkeypatch.setenv('HNO_AXIS_PERM', 'force')
    assert policy((13, 16, 18)) == (1, 2, 0)
    for spatial in ((155, 240, 240), (240, 155, 240), (13, 16, 18), (16, 11, 18)):
        perm = policy(spatial)
        assert sorted(perm) == [0, 1, 2] and spatial[perm[2]] == min(spatial)
